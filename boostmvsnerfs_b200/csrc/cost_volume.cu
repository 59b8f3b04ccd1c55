// K1: fused plane-sweep cost volume (homography warp of S source maps + variance) and the
// depth-hypothesis generators that feed it.  Reference semantics: lib/networks/enerf/utils.py
// :57-95 (homo_warp), :98-153 (get_depth_values), :324-351 (build_feature_volume).
#include <stdlib.h>

#include "bmv_internal.cuh"

namespace bmv {

// view id s of the volume: from DEVICE memory when given (a captured CUDA graph then follows a changed view selection),
// else from the params
__device__ __forceinline__ int view_of(const bmv_cost_volume_params& p, int s) {
  return p.view_dev ? __ldg(p.view_dev + s) : p.view[s];
}

struct WarpTap {
  int off[4];     // element offsets (x,y part) of nw, ne, sw, se taps, clamped in-bounds
  float w[4];     // bilinear weights, 0 for out-of-bounds taps (padding_mode='zeros')
};

// Homography of pixel (x,y) at depth `dep` into one source map, then ATen's bilinear tap set.
__device__ __forceinline__ WarpTap homography_taps(const float* __restrict__ P, float x, float y, float dep,
                                                   int Hs, int Ws, int64_t ys, int64_t xs) {
  // rot @ [x,y,1] accumulated like the reference's bmm, translation / depth added separately
  float cx = add_rn(dot3_gemm(P[0], P[1], P[2], x, y, 1.f), div_rn(P[3], dep));
  float cy = add_rn(dot3_gemm(P[4], P[5], P[6], x, y, 1.f), div_rn(P[7], dep));
  float cz = add_rn(dot3_gemm(P[8], P[9], P[10], x, y, 1.f), div_rn(P[11], dep));
  float zc = fmaxf(cz, 1e-6f);
  if (cz != cz) zc = cz;  // clamp_min propagates NaN
  float u = div_rn(cx, zc), v = div_rn(cy, zc);
  float gx = sub_rn(div_rn(u, (float)(Ws - 1) / 2.f), 1.f);
  float gy = sub_rn(div_rn(v, (float)(Hs - 1) / 2.f), 1.f);
  float ix = unnormalize_ac(gx, Ws), iy = unnormalize_ac(gy, Hs);
  WarpTap t;
  if (!(coord_ok(ix) && coord_ok(iy))) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { t.off[i] = 0; t.w[i] = 0.f; }
    return t;
  }
  float x0 = floorf(ix), y0 = floorf(iy);
  float x1 = x0 + 1.f, y1 = y0 + 1.f;
  float wx1 = ix - x0, wx0 = x1 - ix, wy1 = iy - y0, wy0 = y1 - iy;
  bool vx0 = x0 >= 0.f && x0 <= (float)(Ws - 1), vx1 = x1 >= 0.f && x1 <= (float)(Ws - 1);
  bool vy0 = y0 >= 0.f && y0 <= (float)(Hs - 1), vy1 = y1 >= 0.f && y1 <= (float)(Hs - 1);
  int ix0 = min(max((int)x0, 0), Ws - 1), ix1 = min(max((int)x1, 0), Ws - 1);
  int iy0 = min(max((int)y0, 0), Hs - 1), iy1 = min(max((int)y1, 0), Hs - 1);
  t.off[0] = (int)(iy0 * ys + ix0 * xs); t.w[0] = (vx0 && vy0) ? wx0 * wy0 : 0.f;
  t.off[1] = (int)(iy0 * ys + ix1 * xs); t.w[1] = (vx1 && vy0) ? wx1 * wy0 : 0.f;
  t.off[2] = (int)(iy1 * ys + ix0 * xs); t.w[2] = (vx0 && vy1) ? wx0 * wy1 : 0.f;
  t.off[3] = (int)(iy1 * ys + ix1 * xs); t.w[3] = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return t;
}

template <typename OutT>
__device__ __forceinline__ void store_out(OutT* p, float v);
template <>
__device__ __forceinline__ void store_out<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void store_out<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
template <>
__device__ __forceinline__ void store_out<__half>(__half* p, float v) { *p = half_sat(v); }

// two fp32 values -> one 32-bit word of 16-bit outputs
template <typename OutT>
__device__ __forceinline__ uint32_t pack_out2(float a, float b);
template <>
__device__ __forceinline__ uint32_t pack_out2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack_out2<__half>(float a, float b) {
  return pack_half2_sat(a, b);
}
template <>
__device__ __forceinline__ uint32_t pack_out2<float>(float a, float b) { return 0u; }   // never used (4-byte branch)

// v1: one thread per voxel (flat index over D*h*w, x fastest -> coalesced planar stores),
// S views x 4 taps gathered per channel straight from L2/L1.
template <int S, typename OutT>
__global__ void __launch_bounds__(256) cost_volume_var_kernel(bmv_cost_volume_params p) {
  __shared__ float sP[S * 12];
  if (threadIdx.x < S * 12) sP[threadIdx.x] = p.proj[view_of(p, threadIdx.x / 12) * 12 + threadIdx.x % 12];
  __syncthreads();
  const int64_t nvox = (int64_t)p.D * p.h * p.w;
  const int64_t vox = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vox >= nvox) return;
  const int x = (int)(vox % p.w);
  const int y = (int)((vox / p.w) % p.h);
  const int d = (int)(vox / ((int64_t)p.w * p.h));
  const float dep = __ldg(p.planes + (int64_t)d * p.planes_d_stride + ((int64_t)y * p.w + x) * p.planes_pix_stride);

  WarpTap tap[S];
  const float* base[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    tap[s] = homography_taps(sP + s * 12, (float)x, (float)y, dep, p.Hs, p.Ws, p.feat_y_stride, p.feat_x_stride);
    base[s] = p.feat + (int64_t)view_of(p, s) * p.feat_view_stride;
  }
  OutT* out = reinterpret_cast<OutT*>(p.out) + (int64_t)d * p.out_d_stride + (int64_t)y * p.out_y_stride +
              (int64_t)x * p.out_x_stride;
  const float invS = 1.f;  // division kept exact below
  (void)invS;
#pragma unroll 4
  for (int c = 0; c < p.C; ++c) {
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const float* f = base[s] + (int64_t)c * p.feat_c_stride;
      float v = tap[s].w[0] * __ldg(f + tap[s].off[0]);
      v = fmaf(tap[s].w[1], __ldg(f + tap[s].off[1]), v);
      v = fmaf(tap[s].w[2], __ldg(f + tap[s].off[2]), v);
      v = fmaf(tap[s].w[3], __ldg(f + tap[s].off[3]), v);
      sum = (s == 0) ? v : add_rn(sum, v);
      sq = (s == 0) ? mul_rn(v, v) : add_rn(sq, mul_rn(v, v));
    }
    float mean = div_rn(sum, (float)S);
    float var = sub_rn(div_rn(sq, (float)S), mul_rn(mean, mean));
    store_out<OutT>(out + (int64_t)c * p.out_c_stride, p.out_scale ? var * __ldg(p.out_scale) : var);
  }
}

// v2 (channels-last fast path): feature maps stored [H][W][C] and the volume [D][h][w][C]
// (torch channels_last / channels_last_3d).  CG = C/CPT adjacent lanes share one voxel, each lane
// owns CPT consecutive channels: every bilinear tap is one 16-byte load per lane and the CG lanes
// of a voxel read one contiguous 64..128 B segment; the result is one 16-byte store per lane
// (8-byte for bf16), again contiguous across the voxel's lanes and across x.
template <int S, int CPT, typename OutT>
__global__ void __launch_bounds__(256) cost_volume_var_cl_kernel(bmv_cost_volume_params p, int CG) {
  __shared__ float sP[S * 12];
  if (threadIdx.x < S * 12) sP[threadIdx.x] = p.proj[view_of(p, threadIdx.x / 12) * 12 + threadIdx.x % 12];
  __syncthreads();
  const int64_t nvox = (int64_t)p.D * p.h * p.w;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t vox = gid / CG;
  const int c0 = (int)(gid % CG) * CPT;
  if (vox >= nvox) return;
  const int x = (int)(vox % p.w);
  const int y = (int)((vox / p.w) % p.h);
  const int d = (int)(vox / ((int64_t)p.w * p.h));
  const float dep = __ldg(p.planes + (int64_t)d * p.planes_d_stride + ((int64_t)y * p.w + x) * p.planes_pix_stride);
  float sum[CPT], sq[CPT];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const WarpTap t = homography_taps(sP + s * 12, (float)x, (float)y, dep, p.Hs, p.Ws, p.feat_y_stride, p.feat_x_stride);
    const float* base = p.feat + (int64_t)view_of(p, s) * p.feat_view_stride + c0;
    float v[CPT];
#pragma unroll
    for (int q = 0; q < CPT; q += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(base + t.off[0] + q));
      const float4 b = __ldg(reinterpret_cast<const float4*>(base + t.off[1] + q));
      const float4 c = __ldg(reinterpret_cast<const float4*>(base + t.off[2] + q));
      const float4 e = __ldg(reinterpret_cast<const float4*>(base + t.off[3] + q));
      v[q + 0] = fmaf(t.w[3], e.x, fmaf(t.w[2], c.x, fmaf(t.w[1], b.x, t.w[0] * a.x)));
      v[q + 1] = fmaf(t.w[3], e.y, fmaf(t.w[2], c.y, fmaf(t.w[1], b.y, t.w[0] * a.y)));
      v[q + 2] = fmaf(t.w[3], e.z, fmaf(t.w[2], c.z, fmaf(t.w[1], b.z, t.w[0] * a.z)));
      v[q + 3] = fmaf(t.w[3], e.w, fmaf(t.w[2], c.w, fmaf(t.w[1], b.w, t.w[0] * a.w)));
    }
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      sum[q] = (s == 0) ? v[q] : add_rn(sum[q], v[q]);
      sq[q] = (s == 0) ? mul_rn(v[q], v[q]) : add_rn(sq[q], mul_rn(v[q], v[q]));
    }
  }
  float var[CPT];
#pragma unroll
  for (int q = 0; q < CPT; ++q) {
    const float mean = div_rn(sum[q], (float)S);
    var[q] = sub_rn(div_rn(sq[q], (float)S), mul_rn(mean, mean));
    if (p.out_scale) var[q] *= __ldg(p.out_scale);
  }
  OutT* out = reinterpret_cast<OutT*>(p.out) + (int64_t)d * p.out_d_stride + (int64_t)y * p.out_y_stride +
              (int64_t)x * p.out_x_stride + c0;
  if constexpr (sizeof(OutT) == 4) {
#pragma unroll
    for (int q = 0; q < CPT; q += 4)
      *reinterpret_cast<float4*>(out + q) = make_float4(var[q], var[q + 1], var[q + 2], var[q + 3]);
  } else {
#pragma unroll
    for (int q = 0; q < CPT; q += 4) {
      uint2 pk;
      pk.x = pack_out2<OutT>(var[q], var[q + 1]);
      pk.y = pack_out2<OutT>(var[q + 2], var[q + 3]);
      *reinterpret_cast<uint2*>(out + q) = pk;
    }
  }
}

// v3 (channels-last, the per-frame path).  Two measured problems of v2 (profiles/round1b_*):
//  (a) issue-bound: 1036 instructions per lane, ~75 % of them the op-for-op emulation of the
//      reference's coordinate arithmetic (9 IEEE divisions per view) repeated by every lane of a voxel;
//  (b) 10x more L2->SM traffic than compulsory because consecutive planes of a pixel — which
//      re-touch almost the same source texels — ran in different CTAs.
// v3: one CTA owns VPB consecutive x of ONE row and loops over a group of DG planes, so the sliding
// window of source texels stays in L1; R*[x,y,1] is hoisted out of the plane loop; divisions become
// one correctly-rounded reciprocal each (T/d -> T*rcp(d), x/z -> x*rcp(z), /((W-1)/2) -> *2/(W-1)).
// The coordinate differs from the reference's by <= 2 ulp (~1e-5 px), the same order as the
// reference's own CPU-vs-CUDA difference (ATen multiplies by the reciprocal of scalar divisors on CUDA).
struct FastTap { int off[4]; float w[4]; };

// 4 consecutive channels of a source texel: fp32 maps (16-byte load) or fp16 maps (8-byte load, half the L1 wavefronts)
template <typename FeatT>
__device__ __forceinline__ float4 ld_feat4(const FeatT* q) {
  if constexpr (sizeof(FeatT) == 4) {
    return __ldg(reinterpret_cast<const float4*>(q));
  } else {
    const uint2 r = __ldg(reinterpret_cast<const uint2*>(q));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

__device__ __forceinline__ FastTap fast_taps(float ax, float ay, float az, const float* __restrict__ P, float idep,
                                             float sx, float sy, int Hs, int Ws, int ys, int xs) {
  const float cx = fmaf(P[3], idep, ax), cy = fmaf(P[7], idep, ay), cz = fmaf(P[11], idep, az);
  const float iz = __frcp_rn(fmaxf(cz, 1e-6f));
  // g = u/((W-1)/2) - 1 ; ix = ((g+1)/2)*(W-1)
  const float gx = fmaf(cx * iz, sx, -1.f), gy = fmaf(cy * iz, sy, -1.f);
  const float ix = (gx + 1.f) * 0.5f * (float)(Ws - 1), iy = (gy + 1.f) * 0.5f * (float)(Hs - 1);
  FastTap t;
  const float x0 = floorf(ix), y0 = floorf(iy);
  const float wx1 = ix - x0, wy1 = iy - y0, wx0 = 1.f - wx1, wy0 = 1.f - wy1;
  // NaN / huge coordinates fail every comparison -> all four taps invalid (zeros padding)
  const bool vx0 = x0 >= 0.f && x0 <= (float)(Ws - 1), vx1 = x0 >= -1.f && x0 <= (float)(Ws - 2);
  const bool vy0 = y0 >= 0.f && y0 <= (float)(Hs - 1), vy1 = y0 >= -1.f && y0 <= (float)(Hs - 2);
  const int ix0 = vx0 ? (int)x0 : 0, ix1 = vx1 ? (int)x0 + 1 : 0;
  const int iy0 = vy0 ? (int)y0 : 0, iy1 = vy1 ? (int)y0 + 1 : 0;
  t.off[0] = iy0 * ys + ix0 * xs; t.w[0] = (vx0 && vy0) ? wx0 * wy0 : 0.f;
  t.off[1] = iy0 * ys + ix1 * xs; t.w[1] = (vx1 && vy0) ? wx1 * wy0 : 0.f;
  t.off[2] = iy1 * ys + ix0 * xs; t.w[2] = (vx0 && vy1) ? wx0 * wy1 : 0.f;
  t.off[3] = iy1 * ys + ix1 * xs; t.w[3] = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return t;
}

template <int S, int CPT, typename OutT>
__global__ void __launch_bounds__(256) cost_volume_var_cl3_kernel(bmv_cost_volume_params p, int CG, int DG) {
  __shared__ float sP[S * 12];
  if (threadIdx.x < S * 12) sP[threadIdx.x] = p.proj[view_of(p, threadIdx.x / 12) * 12 + threadIdx.x % 12];
  __syncthreads();
  const int vpb = blockDim.x / CG;                       // voxels (consecutive x) per CTA
  const int x = blockIdx.x * vpb + (int)threadIdx.x / CG;
  const int c0 = ((int)threadIdx.x % CG) * CPT;
  const int y = blockIdx.y;
  if (x >= p.w) return;
  const int d_begin = blockIdx.z * DG, d_end = min(p.D, d_begin + DG);
  const float fx = (float)x, fy = (float)y;
  float ax[S], ay[S], az[S];
  const float* base[S];
#pragma unroll
  for (int s = 0; s < S; ++s) {
    const float* P = sP + s * 12;
    ax[s] = dot3_gemm(P[0], P[1], P[2], fx, fy, 1.f);
    ay[s] = dot3_gemm(P[4], P[5], P[6], fx, fy, 1.f);
    az[s] = dot3_gemm(P[8], P[9], P[10], fx, fy, 1.f);
    base[s] = p.feat + (int64_t)view_of(p, s) * p.feat_view_stride + c0;
  }
  const float sx = 2.f / (float)(p.Ws - 1), sy = 2.f / (float)(p.Hs - 1);
  const int ys = (int)p.feat_y_stride, xs = (int)p.feat_x_stride;
  const float* pl = p.planes + ((int64_t)y * p.w + x) * p.planes_pix_stride;
  OutT* outp = reinterpret_cast<OutT*>(p.out) + (int64_t)y * p.out_y_stride + (int64_t)x * p.out_x_stride + c0;
  for (int d = d_begin; d < d_end; ++d) {
    const float idep = __frcp_rn(__ldg(pl + (int64_t)d * p.planes_d_stride));
    float sum[CPT], sq[CPT];
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const FastTap t = fast_taps(ax[s], ay[s], az[s], sP + s * 12, idep, sx, sy, p.Hs, p.Ws, ys, xs);
#pragma unroll
      for (int q = 0; q < CPT; q += 4) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(base[s] + t.off[0] + q));
        const float4 b = __ldg(reinterpret_cast<const float4*>(base[s] + t.off[1] + q));
        const float4 c = __ldg(reinterpret_cast<const float4*>(base[s] + t.off[2] + q));
        const float4 e = __ldg(reinterpret_cast<const float4*>(base[s] + t.off[3] + q));
        const float v0 = fmaf(t.w[3], e.x, fmaf(t.w[2], c.x, fmaf(t.w[1], b.x, t.w[0] * a.x)));
        const float v1 = fmaf(t.w[3], e.y, fmaf(t.w[2], c.y, fmaf(t.w[1], b.y, t.w[0] * a.y)));
        const float v2 = fmaf(t.w[3], e.z, fmaf(t.w[2], c.z, fmaf(t.w[1], b.z, t.w[0] * a.z)));
        const float v3 = fmaf(t.w[3], e.w, fmaf(t.w[2], c.w, fmaf(t.w[1], b.w, t.w[0] * a.w)));
        if (s == 0) {
          sum[q] = v0; sum[q + 1] = v1; sum[q + 2] = v2; sum[q + 3] = v3;
          sq[q] = v0 * v0; sq[q + 1] = v1 * v1; sq[q + 2] = v2 * v2; sq[q + 3] = v3 * v3;
        } else {
          sum[q] += v0; sum[q + 1] += v1; sum[q + 2] += v2; sum[q + 3] += v3;
          sq[q] = fmaf(v0, v0, sq[q]); sq[q + 1] = fmaf(v1, v1, sq[q + 1]);
          sq[q + 2] = fmaf(v2, v2, sq[q + 2]); sq[q + 3] = fmaf(v3, v3, sq[q + 3]);
        }
      }
    }
    constexpr float invS = 1.f / S;
    float var[CPT];
#pragma unroll
    for (int q = 0; q < CPT; ++q) {
      const float mean = sum[q] * invS;
      var[q] = fmaf(-mean, mean, sq[q] * invS);
      if (p.out_scale) var[q] *= __ldg(p.out_scale);
    }
    OutT* out = outp + (int64_t)d * p.out_d_stride;
    if constexpr (sizeof(OutT) == 4) {
#pragma unroll
      for (int q = 0; q < CPT; q += 4)
        *reinterpret_cast<float4*>(out + q) = make_float4(var[q], var[q + 1], var[q + 2], var[q + 3]);
    } else {
#pragma unroll
      for (int q = 0; q < CPT; q += 4) {
        uint2 pk;
        pk.x = pack_out2<OutT>(var[q], var[q + 1]);
        pk.y = pack_out2<OutT>(var[q + 2], var[q + 3]);
        *reinterpret_cast<uint2*>(out + q) = pk;
      }
    }
  }
}

// v5: as v3, but the bilinear tap set of a (voxel, view, plane) is computed ONCE per warp and handed
// to the channel lanes with warp shuffles (ncu on v3: issue-bound, 406 instructions per lane and
// plane, ~2/3 of them the tap arithmetic repeated by all CG lanes of a voxel; a CTA-level tap table
// in shared memory (v4, two barriers per plane) fixed the instruction count but exposed barrier
// latency on the 8-plane level: 84 us vs 75 us).
// A warp owns VW = 32/CG consecutive x; per round it handles PB planes so that VW*S*PB <= 32 tap
// tasks fill the warp: lane L computes task (plane slot, view, voxel) = (L / (VW*S), (L % (VW*S)) / VW,
// L % VW); the consumer lane of voxel v fetches its 4 offsets + 4 weights per view with 8 shuffles.
// dbg (environment BMV_K1_DEBUG, measurement only — WRONG results): what a staged source tile could buy at most.  Bit 0:
// every tap reads texel 0 of its view (the loads stay, they all hit L1 and coalesce: no scatter, no misses); bit 1: no
// tap loads at all (the instruction floor of everything else).
template <int S, int CG, int PB, typename OutT, typename FeatT = float>
__global__ void __launch_bounds__(256) cost_volume_var_cl5_kernel(bmv_cost_volume_params p, int DG, int dbg) {
  constexpr int VW = 32 / CG;                            // voxels per warp
  constexpr int TASKS = VW * S * PB;
  static_assert(TASKS <= 32, "tap tasks must fit one warp");
  __shared__ float sP[S * 12];
  if (threadIdx.x < S * 12) sP[threadIdx.x] = p.proj[view_of(p, threadIdx.x / 12) * 12 + threadIdx.x % 12];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x_base = (blockIdx.x * 8 + warp) * VW;
  const int y = blockIdx.y;
  if (x_base >= p.w) return;                             // whole warp out of the row
  const int d_begin = blockIdx.z * DG, d_end = min(p.D, d_begin + DG);
  const float sx = 2.f / (float)(p.Ws - 1), sy = 2.f / (float)(p.Hs - 1);
  const int ys = (int)p.feat_y_stride, xs = (int)p.feat_x_stride;
  // ---- tap-task role of this lane
  const bool tapper = lane < TASKS;
  const int t_pl = lane / (VW * S), t_s = (lane % (VW * S)) / VW, t_v = lane % VW;
  const int t_x = min(x_base + t_v, p.w - 1);
  const float* tP = sP + (tapper ? t_s : 0) * 12;
  const float ax = dot3_gemm(tP[0], tP[1], tP[2], (float)t_x, (float)y, 1.f);
  const float ay = dot3_gemm(tP[4], tP[5], tP[6], (float)t_x, (float)y, 1.f);
  const float az = dot3_gemm(tP[8], tP[9], tP[10], (float)t_x, (float)y, 1.f);
  const float* t_planes = p.planes + ((int64_t)y * p.w + t_x) * p.planes_pix_stride;
  // ---- consumer role: voxel v of the warp, 4 channels starting at c0
  const int v = lane / CG, c0 = (lane % CG) * 4;
  const int x = x_base + v;
  const bool active = x < p.w;
  const FeatT* base[S];
#pragma unroll
  for (int s = 0; s < S; ++s) base[s] = reinterpret_cast<const FeatT*>(p.feat) + (int64_t)view_of(p, s) * p.feat_view_stride + c0;
  OutT* outp = reinterpret_cast<OutT*>(p.out) + (int64_t)y * p.out_y_stride + (int64_t)min(x, p.w - 1) * p.out_x_stride + c0;
  constexpr float invS = 1.f / S;
  const float osc = p.out_scale ? __ldg(p.out_scale) : 1.f;
  for (int d0 = d_begin; d0 < d_end; d0 += PB) {
    FastTap t;
    {
      const int d = min(d0 + t_pl, d_end - 1);
      const float idep = __frcp_rn(__ldg(t_planes + (int64_t)d * p.planes_d_stride));
      t = fast_taps(ax, ay, az, tP, idep, sx, sy, p.Hs, p.Ws, ys, xs);
    }
#pragma unroll
    for (int pl = 0; pl < PB; ++pl) {
      const int d = d0 + pl;
      float4 sum, sq;
#pragma unroll
      for (int s = 0; s < S; ++s) {
        const int src = pl * VW * S + s * VW + v;
        const int o0 = __shfl_sync(0xffffffffu, t.off[0], src), o1 = __shfl_sync(0xffffffffu, t.off[1], src);
        const int o2 = __shfl_sync(0xffffffffu, t.off[2], src), o3 = __shfl_sync(0xffffffffu, t.off[3], src);
        const float w0 = __shfl_sync(0xffffffffu, t.w[0], src), w1 = __shfl_sync(0xffffffffu, t.w[1], src);
        const float w2 = __shfl_sync(0xffffffffu, t.w[2], src), w3 = __shfl_sync(0xffffffffu, t.w[3], src);
        float4 a, b, c, e;
        if (dbg & 2) {
          a = make_float4(w0, w1, w2, w3); b = a; c = a; e = a;
        } else {
          const int m = (dbg & 1) ? 0 : -1;
          a = ld_feat4(base[s] + (o0 & m)); b = ld_feat4(base[s] + (o1 & m)); c = ld_feat4(base[s] + (o2 & m)); e = ld_feat4(base[s] + (o3 & m));
        }
        const float v0 = fmaf(w3, e.x, fmaf(w2, c.x, fmaf(w1, b.x, w0 * a.x)));
        const float v1 = fmaf(w3, e.y, fmaf(w2, c.y, fmaf(w1, b.y, w0 * a.y)));
        const float v2 = fmaf(w3, e.z, fmaf(w2, c.z, fmaf(w1, b.z, w0 * a.z)));
        const float v3 = fmaf(w3, e.w, fmaf(w2, c.w, fmaf(w1, b.w, w0 * a.w)));
        if (s == 0) {
          sum = make_float4(v0, v1, v2, v3);
          sq = make_float4(v0 * v0, v1 * v1, v2 * v2, v3 * v3);
        } else {
          sum.x += v0; sum.y += v1; sum.z += v2; sum.w += v3;
          sq.x = fmaf(v0, v0, sq.x); sq.y = fmaf(v1, v1, sq.y); sq.z = fmaf(v2, v2, sq.z); sq.w = fmaf(v3, v3, sq.w);
        }
      }
      if (active && d < d_end) {
        float4 var;
        { const float m = sum.x * invS; var.x = fmaf(-m, m, sq.x * invS); }
        { const float m = sum.y * invS; var.y = fmaf(-m, m, sq.y * invS); }
        { const float m = sum.z * invS; var.z = fmaf(-m, m, sq.z * invS); }
        { const float m = sum.w * invS; var.w = fmaf(-m, m, sq.w * invS); }
        var.x *= osc; var.y *= osc; var.z *= osc; var.w *= osc;
        OutT* out = outp + (int64_t)d * p.out_d_stride;
        if constexpr (sizeof(OutT) == 4) {
          *reinterpret_cast<float4*>(out) = var;
        } else {
          uint2 pk;
          pk.x = pack_out2<OutT>(var.x, var.y);
          pk.y = pack_out2<OutT>(var.z, var.w);
          *reinterpret_cast<uint2*>(out) = pk;
        }
      }
    }
  }
}

// 8 consecutive channels of a source texel: ONE 256-bit load for fp32 maps (LDG.E.ENL2.256, 32-byte aligned), one
// 128-bit load for fp16 maps
template <typename FeatT>
__device__ __forceinline__ void ld_feat8(const FeatT* q, float4& lo, float4& hi) {
  if constexpr (sizeof(FeatT) == 4) {
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y), "=f"(hi.z), "=f"(hi.w) : "l"(q));
  } else {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(q));
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&r.z)), d = __half22float2(*reinterpret_cast<const __half2*>(&r.w));
    lo = make_float4(a.x, a.y, b.x, b.y);
    hi = make_float4(c.x, c.y, d.x, d.y);
  }
}

// bilinear blend of 4 channels (the FMA order of every K1 generation: w0*a, +w1*b, +w2*c, +w3*e)
__device__ __forceinline__ float4 blend4(float w0, float w1, float w2, float w3, const float4& a, const float4& b, const float4& c,
                                         const float4& e) {
  return make_float4(fmaf(w3, e.x, fmaf(w2, c.x, fmaf(w1, b.x, w0 * a.x))), fmaf(w3, e.y, fmaf(w2, c.y, fmaf(w1, b.y, w0 * a.y))),
                     fmaf(w3, e.z, fmaf(w2, c.z, fmaf(w1, b.z, w0 * a.z))), fmaf(w3, e.w, fmaf(w2, c.w, fmaf(w1, b.w, w0 * a.w))));
}
__device__ __forceinline__ void acc_first(float4& sum, float4& sq, const float4& v) {
  sum = v;
  sq = make_float4(v.x * v.x, v.y * v.y, v.z * v.z, v.w * v.w);
}
__device__ __forceinline__ void acc_next(float4& sum, float4& sq, const float4& v) {
  sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
  sq.x = fmaf(v.x, v.x, sq.x); sq.y = fmaf(v.y, v.y, sq.y); sq.z = fmaf(v.z, v.z, sq.z); sq.w = fmaf(v.w, v.w, sq.w);
}
__device__ __forceinline__ float4 variance4(const float4& sum, const float4& sq, float invS, float osc) {
  float4 var;
  { const float m = sum.x * invS; var.x = fmaf(-m, m, sq.x * invS) * osc; }
  { const float m = sum.y * invS; var.y = fmaf(-m, m, sq.y * invS) * osc; }
  { const float m = sum.z * invS; var.z = fmaf(-m, m, sq.z * invS) * osc; }
  { const float m = sum.w * invS; var.w = fmaf(-m, m, sq.w * invS) * osc; }
  return var;
}
// 8 output channels of a voxel: one 32-byte (fp32) / 16-byte (16-bit) store
template <typename OutT>
__device__ __forceinline__ void store_var8(OutT* out, const float4& lo, const float4& hi) {
  if constexpr (sizeof(OutT) == 4) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(out), "f"(lo.x), "f"(lo.y), "f"(lo.z), "f"(lo.w),
                 "f"(hi.x), "f"(hi.y), "f"(hi.z), "f"(hi.w) : "memory");
  } else {
    uint4 pk;
    pk.x = pack_out2<OutT>(lo.x, lo.y); pk.y = pack_out2<OutT>(lo.z, lo.w);
    pk.z = pack_out2<OutT>(hi.x, hi.y); pk.w = pack_out2<OutT>(hi.z, hi.w);
    *reinterpret_cast<uint4*>(out) = pk;
  }
}

// v6: v5 with EIGHT channels per lane.  ncu on v5 (profiles/round2t_k1_staging_whatif.md): the kernel is bound by its
// ~1.2 k thread instructions per voxel, most of them per-LANE overhead (the warp-wide tap computation, 8 shuffles and 4
// address computations per view) that 4 channels of payload amortise badly.  With 8 channels per lane a warp covers
// VW = 32 / (C / 8) voxels (16 at C = 16, 8 at C = 32), every tap is one 256-bit load and the result one 16- / 32-byte
// store: the same loads, blends and stores per voxel, half the overhead.  VW * S tap tasks no longer fit 32 lanes: the
// lanes compute them in NP passes (task T = pass * 32 + lane = (view T / VW, voxel T % VW); VW divides 32, so the voxel
// of a lane's tasks — and its depth hypothesis — is the same in every pass, and the pass of (view s, voxel v) is the
// compile-time constant s * VW / 32).  Same arithmetic per channel as v5: bit-identical volumes (tests).
template <int S, int CG, typename OutT, typename FeatT = float>
__global__ void __launch_bounds__(256) cost_volume_var_cl6_kernel(bmv_cost_volume_params p, int DG) {
  constexpr int VW = 32 / CG;                            // voxels per warp
  constexpr int NP = (VW * S + 31) / 32;                 // tap passes per plane
  __shared__ float sP[S * 12];
  if (threadIdx.x < S * 12) sP[threadIdx.x] = p.proj[view_of(p, threadIdx.x / 12) * 12 + threadIdx.x % 12];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x_base = (blockIdx.x * 8 + warp) * VW;
  const int y = blockIdx.y;
  if (x_base >= p.w) return;                             // whole warp out of the row
  const int d_begin = blockIdx.z * DG, d_end = min(p.D, d_begin + DG);
  const float sx = 2.f / (float)(p.Ws - 1), sy = 2.f / (float)(p.Hs - 1);
  const int ys = (int)p.feat_y_stride, xs = (int)p.feat_x_stride;
  // ---- tap-task roles of this lane
  const int t_v = lane % VW;
  const int t_x = min(x_base + t_v, p.w - 1);
  const float* tP[NP];
  float ax[NP], ay[NP], az[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int t_s = min((q * 32 + lane) / VW, S - 1);
    tP[q] = sP + t_s * 12;
    ax[q] = dot3_gemm(tP[q][0], tP[q][1], tP[q][2], (float)t_x, (float)y, 1.f);
    ay[q] = dot3_gemm(tP[q][4], tP[q][5], tP[q][6], (float)t_x, (float)y, 1.f);
    az[q] = dot3_gemm(tP[q][8], tP[q][9], tP[q][10], (float)t_x, (float)y, 1.f);
  }
  const float* t_planes = p.planes + ((int64_t)y * p.w + t_x) * p.planes_pix_stride;
  // ---- consumer role: voxel v of the warp, 8 channels starting at c0
  const int v = lane / CG, c0 = (lane % CG) * 8;
  const int x = x_base + v;
  const bool active = x < p.w;
  const FeatT* base[S];
#pragma unroll
  for (int s = 0; s < S; ++s) base[s] = reinterpret_cast<const FeatT*>(p.feat) + (int64_t)view_of(p, s) * p.feat_view_stride + c0;
  OutT* outp = reinterpret_cast<OutT*>(p.out) + (int64_t)y * p.out_y_stride + (int64_t)min(x, p.w - 1) * p.out_x_stride + c0;
  constexpr float invS = 1.f / S;
  const float osc = p.out_scale ? __ldg(p.out_scale) : 1.f;
  // the hypothesis of plane d + 1 is loaded while plane d is processed: hypothesis -> taps -> texel loads would otherwise
  // be two memory latencies in a row per plane (ncu: 4.5 long-scoreboard stalls per issue)
  float dep_next = __ldg(t_planes + (int64_t)d_begin * p.planes_d_stride);
  for (int d = d_begin; d < d_end; ++d) {
    const float idep = __frcp_rn(dep_next);
    dep_next = __ldg(t_planes + (int64_t)min(d + 1, d_end - 1) * p.planes_d_stride);
    FastTap t[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) t[q] = fast_taps(ax[q], ay[q], az[q], tP[q], idep, sx, sy, p.Hs, p.Ws, ys, xs);
    float4 sum0, sq0, sum1, sq1;
#pragma unroll
    for (int s = 0; s < S; ++s) {
      const int q = (s * VW) / 32;                       // compile-time after unrolling
      const int src = (s * VW) % 32 + v;
      const int o0 = __shfl_sync(0xffffffffu, t[q].off[0], src), o1 = __shfl_sync(0xffffffffu, t[q].off[1], src);
      const int o2 = __shfl_sync(0xffffffffu, t[q].off[2], src), o3 = __shfl_sync(0xffffffffu, t[q].off[3], src);
      const float w0 = __shfl_sync(0xffffffffu, t[q].w[0], src), w1 = __shfl_sync(0xffffffffu, t[q].w[1], src);
      const float w2 = __shfl_sync(0xffffffffu, t[q].w[2], src), w3 = __shfl_sync(0xffffffffu, t[q].w[3], src);
      float4 a0, a1, b0, b1, c0v, c1v, e0, e1;
      ld_feat8(base[s] + o0, a0, a1); ld_feat8(base[s] + o1, b0, b1); ld_feat8(base[s] + o2, c0v, c1v); ld_feat8(base[s] + o3, e0, e1);
      const float4 v0 = blend4(w0, w1, w2, w3, a0, b0, c0v, e0), v1 = blend4(w0, w1, w2, w3, a1, b1, c1v, e1);
      if (s == 0) { acc_first(sum0, sq0, v0); acc_first(sum1, sq1, v1); }
      else { acc_next(sum0, sq0, v0); acc_next(sum1, sq1, v1); }
    }
    if (active) store_var8<OutT>(outp + (int64_t)d * p.out_d_stride, variance4(sum0, sq0, invS, osc), variance4(sum1, sq1, invS, osc));
  }
}

// v5-multi: the K cost volumes of cascade level 0 in ONE launch.  They share the target frustum and the depth
// hypotheses, and their view triples are drawn from the same N source views, so a warped feature of view u at
// (voxel, plane) is identical in every chain that contains u: gather each of the U UNIQUE views once (U x 4 taps
// instead of K x S x 4 — 24 instead of 48 for K = 4 triples out of 6 views) and keep one (sum, sum of squares) pair per
// chain.  Same tap sharing through warp shuffles as v5; PB = 1.
constexpr int kMultiMaxK = 4;
template <int CG, typename OutT, typename FeatT = float>
__global__ void __launch_bounds__(256) cost_volume_var_multi_kernel(bmv_cost_volume_multi_params mp, int DG) {
  const bmv_cost_volume_params& p = mp.b;
  constexpr int VW = 32 / CG;                            // voxels per warp
  const int U = p.S, K = mp.K;
  __shared__ float sP[BMV_MAX_VIEWS * 12];
  __shared__ int s_uview[BMV_MAX_VIEWS], s_mask[BMV_MAX_VIEWS];
  if (threadIdx.x < U) {
    int vw = p.view[threadIdx.x], mask = mp.chain_mask[threadIdx.x];
    if (mp.triples_dev) {                                // unique view u IS source view u; masks from the device table
      vw = threadIdx.x;
      mask = 0;
      for (int k = 0; k < K; ++k)
        for (int j = 0; j < mp.views_per_chain; ++j)
          if (__ldg(mp.triples_dev + k * mp.views_per_chain + j) == vw) mask |= 1 << k;
    }
    s_uview[threadIdx.x] = vw;
    s_mask[threadIdx.x] = mask;
  }
  __syncthreads();
  if (threadIdx.x < U * 12) sP[threadIdx.x] = p.proj[s_uview[threadIdx.x / 12] * 12 + threadIdx.x % 12];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x_base = (blockIdx.x * 8 + warp) * VW;
  const int y = blockIdx.y;
  if (x_base >= p.w) return;
  const int d_begin = blockIdx.z * DG, d_end = min(p.D, d_begin + DG);
  const float sx = 2.f / (float)(p.Ws - 1), sy = 2.f / (float)(p.Hs - 1);
  const int ys = (int)p.feat_y_stride, xs = (int)p.feat_x_stride;
  // ---- tap-task role: lane L computes the taps of (unique view L / VW, voxel L % VW)
  const bool tapper = lane < VW * U;
  const int t_u = tapper ? lane / VW : 0, t_v = lane % VW;
  const int t_x = min(x_base + t_v, p.w - 1);
  const float* tP = sP + t_u * 12;
  const float ax = dot3_gemm(tP[0], tP[1], tP[2], (float)t_x, (float)y, 1.f);
  const float ay = dot3_gemm(tP[4], tP[5], tP[6], (float)t_x, (float)y, 1.f);
  const float az = dot3_gemm(tP[8], tP[9], tP[10], (float)t_x, (float)y, 1.f);
  const float* t_planes = p.planes + ((int64_t)y * p.w + t_x) * p.planes_pix_stride;
  // ---- consumer role: voxel v of the warp, 4 channels starting at c0
  const int v = lane / CG, c0 = (lane % CG) * 4;
  const int x = x_base + v;
  const bool active = x < p.w;
  OutT* outp = reinterpret_cast<OutT*>(p.out) + (int64_t)y * p.out_y_stride + (int64_t)min(x, p.w - 1) * p.out_x_stride + c0;
  const float invS = 1.f / (float)mp.views_per_chain;
  const float osc = p.out_scale ? __ldg(p.out_scale) : 1.f;
  for (int d = d_begin; d < d_end; ++d) {
    const float idep = __frcp_rn(__ldg(t_planes + (int64_t)d * p.planes_d_stride));
    const FastTap t = fast_taps(ax, ay, az, tP, idep, sx, sy, p.Hs, p.Ws, ys, xs);
    float4 sum[kMultiMaxK], sq[kMultiMaxK];
#pragma unroll
    for (int k = 0; k < kMultiMaxK; ++k) { sum[k] = make_float4(0.f, 0.f, 0.f, 0.f); sq[k] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int u = 0; u < BMV_MAX_VIEWS; ++u) {
      if (u < U) {                                       // uniform
        const int src = u * VW + v;
        const int o0 = __shfl_sync(0xffffffffu, t.off[0], src), o1 = __shfl_sync(0xffffffffu, t.off[1], src);
        const int o2 = __shfl_sync(0xffffffffu, t.off[2], src), o3 = __shfl_sync(0xffffffffu, t.off[3], src);
        const float w0 = __shfl_sync(0xffffffffu, t.w[0], src), w1 = __shfl_sync(0xffffffffu, t.w[1], src);
        const float w2 = __shfl_sync(0xffffffffu, t.w[2], src), w3 = __shfl_sync(0xffffffffu, t.w[3], src);
        const int mask = s_mask[u];
        if (mask == 0) continue;                         // view in no chain (device-resident selection): uniform
        const FeatT* base = reinterpret_cast<const FeatT*>(p.feat) + (int64_t)s_uview[u] * p.feat_view_stride + c0;
        const float4 a = ld_feat4(base + o0), b = ld_feat4(base + o1), c = ld_feat4(base + o2), e = ld_feat4(base + o3);
        const float v0 = fmaf(w3, e.x, fmaf(w2, c.x, fmaf(w1, b.x, w0 * a.x)));
        const float v1 = fmaf(w3, e.y, fmaf(w2, c.y, fmaf(w1, b.y, w0 * a.y)));
        const float v2 = fmaf(w3, e.z, fmaf(w2, c.z, fmaf(w1, b.z, w0 * a.z)));
        const float v3 = fmaf(w3, e.w, fmaf(w2, c.w, fmaf(w1, b.w, w0 * a.w)));
#pragma unroll
        for (int k = 0; k < kMultiMaxK; ++k)
          if ((mask >> k) & 1) {                         // uniform
            sum[k].x += v0; sum[k].y += v1; sum[k].z += v2; sum[k].w += v3;
            sq[k].x = fmaf(v0, v0, sq[k].x); sq[k].y = fmaf(v1, v1, sq[k].y);
            sq[k].z = fmaf(v2, v2, sq[k].z); sq[k].w = fmaf(v3, v3, sq[k].w);
          }
      }
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < kMultiMaxK; ++k)
        if (k < K) {
          float4 var;
          { const float m = sum[k].x * invS; var.x = fmaf(-m, m, sq[k].x * invS); }
          { const float m = sum[k].y * invS; var.y = fmaf(-m, m, sq[k].y * invS); }
          { const float m = sum[k].z * invS; var.z = fmaf(-m, m, sq[k].z * invS); }
          { const float m = sum[k].w * invS; var.w = fmaf(-m, m, sq[k].w * invS); }
          var.x *= osc; var.y *= osc; var.z *= osc; var.w *= osc;
          OutT* out = outp + (int64_t)k * mp.out_k_stride + (int64_t)d * p.out_d_stride;
          if constexpr (sizeof(OutT) == 4) {
            *reinterpret_cast<float4*>(out) = var;
          } else {
            uint2 pk;
            pk.x = pack_out2<OutT>(var.x, var.y);
            pk.y = pack_out2<OutT>(var.z, var.w);
            *reinterpret_cast<uint2*>(out) = pk;
          }
        }
    }
  }
}

// v6-multi: v5-multi with eight channels per lane (see v6): VW = 32 / (C / 8) voxels per warp, the VW * U tap tasks
// computed in up to two passes, 2 x K (sum, sum of squares) float4 pairs per lane.
template <int CG, typename OutT, typename FeatT = float>
__global__ void __launch_bounds__(256, 2) cost_volume_var_multi6_kernel(bmv_cost_volume_multi_params mp, int DG) {
  const bmv_cost_volume_params& p = mp.b;
  constexpr int VW = 32 / CG;                            // voxels per warp
  constexpr int MAXU = 64 / VW;                          // unique views: the VW * U tap tasks take two passes at most
  constexpr int NP = 2;
  const int U = p.S, K = mp.K;
  __shared__ float sP[BMV_MAX_VIEWS * 12];
  __shared__ int s_uview[BMV_MAX_VIEWS], s_mask[BMV_MAX_VIEWS];
  if (threadIdx.x < U) {
    int vw = p.view[threadIdx.x], mask = mp.chain_mask[threadIdx.x];
    if (mp.triples_dev) {                                // unique view u IS source view u; masks from the device table
      vw = threadIdx.x;
      mask = 0;
      for (int k = 0; k < K; ++k)
        for (int j = 0; j < mp.views_per_chain; ++j)
          if (__ldg(mp.triples_dev + k * mp.views_per_chain + j) == vw) mask |= 1 << k;
    }
    s_uview[threadIdx.x] = vw;
    s_mask[threadIdx.x] = mask;
  }
  __syncthreads();
  if (threadIdx.x < U * 12) sP[threadIdx.x] = p.proj[s_uview[threadIdx.x / 12] * 12 + threadIdx.x % 12];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int x_base = (blockIdx.x * 8 + warp) * VW;
  const int y = blockIdx.y;
  if (x_base >= p.w) return;
  const int d_begin = blockIdx.z * DG, d_end = min(p.D, d_begin + DG);
  const float sx = 2.f / (float)(p.Ws - 1), sy = 2.f / (float)(p.Hs - 1);
  const int ys = (int)p.feat_y_stride, xs = (int)p.feat_x_stride;
  // ---- tap-task roles: in pass q lane L computes the taps of (unique view (q * 32 + L) / VW, voxel L % VW)
  const int t_v = lane % VW;
  const int t_x = min(x_base + t_v, p.w - 1);
  const float* tP[NP];
  float ax[NP], ay[NP], az[NP];
#pragma unroll
  for (int q = 0; q < NP; ++q) {
    const int t_u = min((q * 32 + lane) / VW, U - 1);
    tP[q] = sP + t_u * 12;
    ax[q] = dot3_gemm(tP[q][0], tP[q][1], tP[q][2], (float)t_x, (float)y, 1.f);
    ay[q] = dot3_gemm(tP[q][4], tP[q][5], tP[q][6], (float)t_x, (float)y, 1.f);
    az[q] = dot3_gemm(tP[q][8], tP[q][9], tP[q][10], (float)t_x, (float)y, 1.f);
  }
  const bool pass1 = NP > 1 && VW * U > 32;              // uniform
  const float* t_planes = p.planes + ((int64_t)y * p.w + t_x) * p.planes_pix_stride;
  // ---- consumer role: voxel v of the warp, 8 channels starting at c0
  const int v = lane / CG, c0 = (lane % CG) * 8;
  const int x = x_base + v;
  const bool active = x < p.w;
  OutT* outp = reinterpret_cast<OutT*>(p.out) + (int64_t)y * p.out_y_stride + (int64_t)min(x, p.w - 1) * p.out_x_stride + c0;
  const float invS = 1.f / (float)mp.views_per_chain;
  const float osc = p.out_scale ? __ldg(p.out_scale) : 1.f;
  float dep_next = __ldg(t_planes + (int64_t)d_begin * p.planes_d_stride);
  for (int d = d_begin; d < d_end; ++d) {
    const float idep = __frcp_rn(dep_next);
    dep_next = __ldg(t_planes + (int64_t)min(d + 1, d_end - 1) * p.planes_d_stride);
    FastTap t[NP];
    t[0] = fast_taps(ax[0], ay[0], az[0], tP[0], idep, sx, sy, p.Hs, p.Ws, ys, xs);
    if (NP > 1) {
      if (pass1) t[NP - 1] = fast_taps(ax[NP - 1], ay[NP - 1], az[NP - 1], tP[NP - 1], idep, sx, sy, p.Hs, p.Ws, ys, xs);
      else t[NP - 1] = t[0];
    }
    float4 sum[kMultiMaxK][2], sq[kMultiMaxK][2];
#pragma unroll
    for (int k = 0; k < kMultiMaxK; ++k)
#pragma unroll
      for (int h = 0; h < 2; ++h) { sum[k][h] = make_float4(0.f, 0.f, 0.f, 0.f); sq[k][h] = make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
      if (u < U) {                                       // uniform
        const int q = (u * VW) / 32;                     // compile-time after unrolling
        const int src = (u * VW) % 32 + v;
        const int o0 = __shfl_sync(0xffffffffu, t[q].off[0], src), o1 = __shfl_sync(0xffffffffu, t[q].off[1], src);
        const int o2 = __shfl_sync(0xffffffffu, t[q].off[2], src), o3 = __shfl_sync(0xffffffffu, t[q].off[3], src);
        const float w0 = __shfl_sync(0xffffffffu, t[q].w[0], src), w1 = __shfl_sync(0xffffffffu, t[q].w[1], src);
        const float w2 = __shfl_sync(0xffffffffu, t[q].w[2], src), w3 = __shfl_sync(0xffffffffu, t[q].w[3], src);
        const int mask = s_mask[u];
        if (mask == 0) continue;                         // view in no chain (device-resident selection): uniform
        const FeatT* base = reinterpret_cast<const FeatT*>(p.feat) + (int64_t)s_uview[u] * p.feat_view_stride + c0;
        float4 a0, a1, b0, b1, c0v, c1v, e0, e1;
        ld_feat8(base + o0, a0, a1); ld_feat8(base + o1, b0, b1); ld_feat8(base + o2, c0v, c1v); ld_feat8(base + o3, e0, e1);
        const float4 v0 = blend4(w0, w1, w2, w3, a0, b0, c0v, e0), v1 = blend4(w0, w1, w2, w3, a1, b1, c1v, e1);
#pragma unroll
        for (int k = 0; k < kMultiMaxK; ++k)
          if ((mask >> k) & 1) {                         // uniform
            acc_next(sum[k][0], sq[k][0], v0);
            acc_next(sum[k][1], sq[k][1], v1);
          }
      }
    }
    if (active) {
#pragma unroll
      for (int k = 0; k < kMultiMaxK; ++k)
        if (k < K)
          store_var8<OutT>(outp + (int64_t)k * mp.out_k_stride + (int64_t)d * p.out_d_stride, variance4(sum[k][0], sq[k][0], invS, osc),
                           variance4(sum[k][1], sq[k][1], invS, osc));
    }
  }
}

template <typename OutT>
static int launch_cost_volume_multi(const bmv_cost_volume_multi_params& mp, cudaStream_t st) {
  const bmv_cost_volume_params& p = mp.b;
  const int CG = p.C / 4, threads = 256, vpb = threads / CG;
  int DG = p.D;
  while (DG > 2 && (int64_t)p.w * p.h * CG * ((p.D + DG - 1) / DG) < 250000) DG = (DG + 1) / 2;
  // v6 (eight channels per lane) when every voxel's 8-channel group is 32-byte (fp32) / 16-byte (16-bit) addressable
  const int esz = p.feat_half ? 2 : 4;
  static const int variant_env = getenv("BMV_K1_VARIANT") ? atoi(getenv("BMV_K1_VARIANT")) : 0;   // measurements
  const bool wide = p.variant != 5 && variant_env != 5 && p.C % 8 == 0 && (64 / (32 / (p.C / 8))) >= p.S && p.feat_x_stride % 8 == 0 && p.feat_y_stride % 8 == 0 &&
                    p.feat_view_stride % 8 == 0 && ((uintptr_t)p.feat & (8 * esz - 1)) == 0 && p.out_x_stride % 8 == 0 &&
                    p.out_y_stride % 8 == 0 && p.out_d_stride % 8 == 0 && mp.out_k_stride % 8 == 0 &&
                    ((uintptr_t)p.out & (8 * sizeof(OutT) - 1)) == 0;
  if (wide) {
    const int CG6 = p.C / 8, vpb6 = threads / CG6;
    int DG6 = p.D;
    while (DG6 > 2 && (int64_t)p.w * p.h * CG6 * ((p.D + DG6 - 1) / DG6) < 250000) DG6 = (DG6 + 1) / 2;
    static const int dg_env = getenv("BMV_K1_DG") ? atoi(getenv("BMV_K1_DG")) : 0;
    if (dg_env > 0) DG6 = dg_env;
    dim3 grid6((p.w + vpb6 - 1) / vpb6, p.h, (p.D + DG6 - 1) / DG6);
    if (p.feat_half) {
      if (CG6 == 4) cost_volume_var_multi6_kernel<4, OutT, __half><<<grid6, threads, 0, st>>>(mp, DG6);
      else cost_volume_var_multi6_kernel<2, OutT, __half><<<grid6, threads, 0, st>>>(mp, DG6);
    } else if (CG6 == 4) cost_volume_var_multi6_kernel<4, OutT><<<grid6, threads, 0, st>>>(mp, DG6);
    else cost_volume_var_multi6_kernel<2, OutT><<<grid6, threads, 0, st>>>(mp, DG6);
    return check_launch("bmv_cost_volume_var_multi");
  }
  dim3 grid((p.w + vpb - 1) / vpb, p.h, (p.D + DG - 1) / DG);
  if (p.feat_half) {
    if (CG == 8) cost_volume_var_multi_kernel<8, OutT, __half><<<grid, threads, 0, st>>>(mp, DG);
    else cost_volume_var_multi_kernel<4, OutT, __half><<<grid, threads, 0, st>>>(mp, DG);
  } else if (CG == 8) cost_volume_var_multi_kernel<8, OutT><<<grid, threads, 0, st>>>(mp, DG);
  else cost_volume_var_multi_kernel<4, OutT><<<grid, threads, 0, st>>>(mp, DG);
  return check_launch("bmv_cost_volume_var_multi");
}

// ---------------------------------------------------------------- range scale of an fp16 volume (bmv_volume_scale)
template <typename T>
__global__ void __launch_bounds__(256) volume_scale_kernel(bmv_volume_scale_params p) {
  constexpr int VEC = 16 / sizeof(T);
  float m = 0.f;
  const int64_t nvec = p.n / VEC;
  const uint4* src = reinterpret_cast<const uint4*>(p.x);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  // four independent 16-byte loads in flight per thread (a dependent one-load loop is latency bound: 20 us for 25 MB)
  for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < nvec; i0 += 4 * stride) {
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u * stride;
      q[u] = i < nvec ? __ldg(src + i) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if constexpr (sizeof(T) == 4) {
        m = fmaxf(m, fmaxf(fmaxf(fabsf(__uint_as_float(q[u].x)), fabsf(__uint_as_float(q[u].y))),
                           fmaxf(fabsf(__uint_as_float(q[u].z)), fabsf(__uint_as_float(q[u].w)))));
      } else {
        const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
          m = fmaxf(m, fmaxf(fabsf(f.x), fabsf(f.y)));
        }
      }
    }
  }
  m = fminf(m, 3.0e38f);                                 // inf -> finite; NaNs were dropped by fmaxf
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __shared__ float sm[8];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) m = fmaxf(m, sm[i]);
    unsigned* scratch = reinterpret_cast<unsigned*>(p.scale) + 2;
    atomicMax(scratch, __float_as_uint(m));              // non-negative floats order like their bit patterns
    __threadfence();
    if (atomicAdd(scratch + 1, 1u) == gridDim.x - 1) {   // last block: every maximum has been merged
      const float mx = __uint_as_float(atomicMax(scratch, 0u));
      int k = 0;
      if (mx > 0.f) {
        int e;
        frexpf(mx, &e);                                  // mx = f * 2^e, f in [0.5, 1)  ->  mx^2 < 2^(2e)
        int te;
        frexpf(p.target, &te);                           // target = g * 2^te, g in [0.5, 1) -> 2^(te-1) <= target
        k = min(max(te - 1 - 2 * e, -40), 40);
      }
      const float cs = p.consumer_scale != 0.f ? p.consumer_scale : 1.f;
      p.scale[0] = ldexpf(1.f, k);
      p.scale[1] = ldexpf(1.f, -k);
      p.scale[4] = ldexpf(1.f, k) * cs;
      p.scale[5] = 1.f / (ldexpf(1.f, k) * cs);
      scratch[0] = 0u;
      scratch[1] = 0u;
    }
  }
}

// ---------------------------------------------------------------- depth hypotheses, level 0
__global__ void depth_planes_first_kernel(bmv_depth_planes_first_params p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float nr = __ldg(p.near_far), fr = __ldg(p.near_far + 1);
  auto plane = [&](int d) {
    float t = __ldg(p.t + d);
    if (p.depth_inv) {
      float dn = div_rn(1.f, nr), df = div_rn(1.f, fr);
      return div_rn(1.f, add_rn(dn, mul_rn(t, sub_rn(df, dn))));
    }
    return add_rn(nr, mul_rn(sub_rn(fr, nr), t));
  };
  if (i < p.D) p.planes[i] = plane(i);
  const int hw = p.h * p.w;
  if (i < hw) {
    float a = plane(0), b = plane(p.D - 1);
    if (p.depth_inv) { a = div_rn(1.f, fmaxf(a, 1e-6f)); b = div_rn(1.f, fmaxf(b, 1e-6f)); }
    p.near_far_out[i] = a;
    p.near_far_out[hw + i] = b;
  }
}

// ---------------------------------------------------------------- depth hypotheses, level >= 1
__global__ void depth_planes_next_kernel(bmv_depth_planes_next_params p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int hw = p.h * p.w;
  if (i >= hw) return;
  const int x = i % p.w, y = i / p.w;
  UpCoord uy = up_coord(y, p.h0, p.h), ux = up_coord(x, p.w0, p.w);
  const int hw0 = p.h0 * p.w0;
  {
    const int b = blockIdx.y;                                           // chain
    p.depth += b * p.depth_b_stride; p.std += b * p.std_b_stride; p.near_far += b * p.nf_b_stride;
    p.planes += (int64_t)b * p.D * hw; p.near_far_out += (int64_t)b * 2 * hw;
  }
  float dep = up_sample(p.depth, p.w0, uy, ux);
  float sd = up_sample(p.std, p.w0, uy, ux);
  float nf0 = up_sample(p.near_far, p.w0, uy, ux);
  float nf1 = up_sample(p.near_far + hw0, p.w0, uy, ux);
  // previous level is in disparity: [d+s, d-s] clamped into [nf0, nf1], then inverted
  float lo = add_rn(dep, sd), hi = sub_rn(dep, sd);
  lo = lo > nf0 ? nf0 : lo;
  hi = hi < nf1 ? nf1 : hi;
  float nr = div_rn(1.f, lo), fr = div_rn(1.f, hi);
  float first = 0.f, last = 0.f;
  for (int d = 0; d < p.D; ++d) {
    float t = __ldg(p.t + d), v;
    if (p.cur_inv) {
      float dn = div_rn(1.f, nr), df = div_rn(1.f, fr);
      v = div_rn(1.f, add_rn(dn, mul_rn(t, sub_rn(df, dn))));
    } else {
      v = add_rn(nr, mul_rn(t, sub_rn(fr, nr)));
    }
    p.planes[(int64_t)d * hw + i] = v;
    if (d == 0) first = v;
    last = v;
  }
  if (p.cur_inv) { first = div_rn(1.f, fmaxf(first, 1e-6f)); last = div_rn(1.f, fmaxf(last, 1e-6f)); }
  p.near_far_out[i] = first;
  p.near_far_out[hw + i] = last;
}

template <int S, typename OutT>
static int launch_cost_volume_s(const bmv_cost_volume_params& p, cudaStream_t st) {
  const int64_t nvox = (int64_t)p.D * p.h * p.w;
  const int threads = 256;
  // channels-last fast path: unit channel stride on both sides, 16-byte aligned rows
  const bool cl = p.feat_c_stride == 1 && p.out_c_stride == 1 && p.C % 4 == 0 && p.feat_x_stride % 4 == 0 &&
                  p.feat_y_stride % 4 == 0 && p.feat_view_stride % 4 == 0 && ((uintptr_t)p.feat & 15) == 0 &&
                  p.out_x_stride % 4 == 0 && p.out_y_stride % 4 == 0 && p.out_d_stride % 4 == 0 &&
                  ((uintptr_t)p.out & (sizeof(OutT) == 4 ? 15 : 7)) == 0;
  const int cpt = (p.C % 8 == 0 && p.C >= 32) ? 8 : 4;
  const int CG = p.C / cpt;
  const int CG3 = p.C / 4;                               // v3 always uses 4 channels per lane
  const bool v3 = cl && p.exact_coords == 0 && threads % CG3 == 0 && CG3 <= threads && p.h <= 65535 &&
                  (int64_t)p.Hs * p.feat_y_stride < (1ll << 31);
  if (p.feat_half && !v3) {
    set_error("bmv_cost_volume_var: fp16 feature maps need the channels-last fast path (exact_coords = 0, aligned tensors)");
    return BMV_ERR_UNSUPPORTED_SHAPE;
  }
  if (v3) {
    // plane groups as deep as possible (L1 reuse of the sliding texel window along d) while still
    // launching >= ~250k lanes (148 SMs x 2048 threads = 303k resident)
    const int vpb = threads / CG3;
    const int xchunks = (p.w + vpb - 1) / vpb;
    int DG = p.D;
    while (DG > 2 && (int64_t)p.w * p.h * CG3 * ((p.D + DG - 1) / DG) < 250000) DG = (DG + 1) / 2;
    dim3 grid(xchunks, p.h, (p.D + DG - 1) / DG);
    static const int dbg = getenv("BMV_K1_DEBUG") ? atoi(getenv("BMV_K1_DEBUG")) : 0;
    // v6 (eight channels per lane) when every voxel's 8-channel group is 32-byte (fp32) / 16-byte (16-bit) addressable
    const int esz = p.feat_half ? 2 : 4;
    static const int variant_env = getenv("BMV_K1_VARIANT") ? atoi(getenv("BMV_K1_VARIANT")) : 0;   // measurements
    const bool wide = p.variant != 5 && variant_env != 5 && dbg == 0 && (p.C == 16 || p.C == 32) && S <= 4 && p.feat_x_stride % 8 == 0 && p.feat_y_stride % 8 == 0 &&
                      p.feat_view_stride % 8 == 0 && ((uintptr_t)p.feat & (8 * esz - 1)) == 0 && p.out_x_stride % 8 == 0 &&
                      p.out_y_stride % 8 == 0 && p.out_d_stride % 8 == 0 && ((uintptr_t)p.out & (8 * sizeof(OutT) - 1)) == 0;
    if constexpr (S <= 4) {
      if (wide) {
        const int CG6 = p.C / 8, vpb6 = threads / CG6;
        int DG6 = p.D;
        while (DG6 > 2 && (int64_t)p.w * p.h * CG6 * ((p.D + DG6 - 1) / DG6) < 250000) DG6 = (DG6 + 1) / 2;
        static const int dg_env = getenv("BMV_K1_DG") ? atoi(getenv("BMV_K1_DG")) : 0;
        if (dg_env > 0) DG6 = dg_env;
        dim3 grid6((p.w + vpb6 - 1) / vpb6, p.h, (p.D + DG6 - 1) / DG6);
        if (CG6 == 4) {
          if (p.feat_half) cost_volume_var_cl6_kernel<S, 4, OutT, __half><<<grid6, threads, 0, st>>>(p, DG6);
          else cost_volume_var_cl6_kernel<S, 4, OutT><<<grid6, threads, 0, st>>>(p, DG6);
        } else {
          if (p.feat_half) cost_volume_var_cl6_kernel<S, 2, OutT, __half><<<grid6, threads, 0, st>>>(p, DG6);
          else cost_volume_var_cl6_kernel<S, 2, OutT><<<grid6, threads, 0, st>>>(p, DG6);
        }
        return check_launch("bmv_cost_volume_var");
      }
    }
    if constexpr (S <= 4) {
      if (CG3 == 8) {              // C = 32: warp = 4 voxels x 8 lanes, 2 planes per round
        if (p.feat_half) cost_volume_var_cl5_kernel<S, 8, 2, OutT, __half><<<grid, threads, 0, st>>>(p, DG, dbg);
        else cost_volume_var_cl5_kernel<S, 8, 2, OutT><<<grid, threads, 0, st>>>(p, DG, dbg);
        return check_launch("bmv_cost_volume_var");
      }
      if (CG3 == 4) {              // C = 16: warp = 8 voxels x 4 lanes, 1 plane per round
        if (p.feat_half) cost_volume_var_cl5_kernel<S, 4, 1, OutT, __half><<<grid, threads, 0, st>>>(p, DG, dbg);
        else cost_volume_var_cl5_kernel<S, 4, 1, OutT><<<grid, threads, 0, st>>>(p, DG, dbg);
        return check_launch("bmv_cost_volume_var");
      }
    }
    if (p.feat_half) {
      set_error("bmv_cost_volume_var: fp16 feature maps are instantiated for C = 16 / 32 with up to 4 views only");
      return BMV_ERR_UNSUPPORTED_SHAPE;
    }
    cost_volume_var_cl3_kernel<S, 4, OutT><<<grid, threads, 0, st>>>(p, CG3, DG);
  } else if (cl) {
    if (cpt == 8) cost_volume_var_cl_kernel<S, 8, OutT><<<(unsigned)ceil_div64(nvox * CG, threads), threads, 0, st>>>(p, CG);
    else cost_volume_var_cl_kernel<S, 4, OutT><<<(unsigned)ceil_div64(nvox * CG, threads), threads, 0, st>>>(p, CG);
  } else {
    cost_volume_var_kernel<S, OutT><<<(unsigned)ceil_div64(nvox, threads), threads, 0, st>>>(p);
  }
  return check_launch("bmv_cost_volume_var");
}

template <typename OutT>
static int launch_cost_volume(const bmv_cost_volume_params& p, cudaStream_t st) {
  switch (p.S) {
    case 1: return launch_cost_volume_s<1, OutT>(p, st);
    case 2: return launch_cost_volume_s<2, OutT>(p, st);
    case 3: return launch_cost_volume_s<3, OutT>(p, st);
    case 4: return launch_cost_volume_s<4, OutT>(p, st);
    default:
      set_error("bmv_cost_volume_var: S=%d views per volume not supported (1..4)", p.S);
      return BMV_ERR_UNSUPPORTED_SHAPE;
  }
}

}  // namespace bmv

extern "C" BMV_API int bmv_cost_volume_var(const bmv_cost_volume_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_cost_volume_var");
  using namespace bmv;
  BMV_REQUIRE(p != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var: null params");
  BMV_REQUIRE(p->feat && p->proj && p->planes && p->out, BMV_ERR_INVALID_ARGUMENT,
              "bmv_cost_volume_var: null device pointer");
  BMV_REQUIRE(p->S >= 1 && p->S <= BMV_MAX_VIEWS && p->C >= 1 && p->Hs >= 1 && p->Ws >= 1 && p->D >= 1 &&
                  p->h >= 1 && p->w >= 1,
              BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var: non-positive size");
  for (int s = 0; s < p->S; ++s)
    BMV_REQUIRE(p->view[s] >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var: negative view index");
  // tap offsets are kept in 32-bit
  BMV_REQUIRE((int64_t)p->Hs * llabs(p->feat_y_stride) + (int64_t)p->Ws * llabs(p->feat_x_stride) < (1ll << 31),
              BMV_ERR_UNSUPPORTED_SHAPE, "bmv_cost_volume_var: source map too large for 32-bit tap offsets");
  cudaStream_t st = (cudaStream_t)stream;
  BMV_REQUIRE(p->out_bf16 >= 0 && p->out_bf16 <= 2, BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var: out_bf16 must be 0, 1 or 2");
  if (p->out_bf16 == 1) return launch_cost_volume<__nv_bfloat16>(*p, st);
  if (p->out_bf16 == 2) return launch_cost_volume<__half>(*p, st);
  return launch_cost_volume<float>(*p, st);
}

extern "C" BMV_API int bmv_cost_volume_var_multi(const bmv_cost_volume_multi_params* mp, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_cost_volume_var_multi");
  using namespace bmv;
  BMV_REQUIRE(mp != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var_multi: null params");
  const bmv_cost_volume_params* p = &mp->b;
  BMV_REQUIRE(p->feat && p->proj && p->planes && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var_multi: null device pointer");
  BMV_REQUIRE(p->S >= 1 && p->S <= BMV_MAX_VIEWS && p->Hs >= 2 && p->Ws >= 2 && p->D >= 1 && p->h >= 1 && p->w >= 1 && p->h <= 65535,
              BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var_multi: bad size");
  BMV_REQUIRE(mp->K >= 1 && mp->K <= kMultiMaxK && mp->views_per_chain >= 1, BMV_ERR_UNSUPPORTED_SHAPE,
              "bmv_cost_volume_var_multi: K=%d chains not supported (1..%d)", mp->K, kMultiMaxK);
  BMV_REQUIRE((p->C == 32 || p->C == 16) && (32 / (p->C / 4)) * p->S <= 32, BMV_ERR_UNSUPPORTED_SHAPE,
              "bmv_cost_volume_var_multi: C=%d with %d unique views not instantiated (C 32: <= 8 views, C 16: <= 4)", p->C, p->S);
  for (int s = 0; s < p->S; ++s)
    BMV_REQUIRE(p->view[s] >= 0 && mp->chain_mask[s] >= 0 && mp->chain_mask[s] < (1 << mp->K), BMV_ERR_INVALID_ARGUMENT,
                "bmv_cost_volume_var_multi: bad view / chain mask");
  BMV_REQUIRE(p->feat_c_stride == 1 && p->out_c_stride == 1 && p->feat_x_stride % 4 == 0 && p->feat_y_stride % 4 == 0 &&
                  p->feat_view_stride % 4 == 0 && ((uintptr_t)p->feat & 15) == 0 && p->out_x_stride % 4 == 0 &&
                  p->out_y_stride % 4 == 0 && p->out_d_stride % 4 == 0 && mp->out_k_stride % 4 == 0 &&
                  ((uintptr_t)p->out & 15) == 0 && p->exact_coords == 0,
              BMV_ERR_UNSUPPORTED_SHAPE, "bmv_cost_volume_var_multi: channels-last, 16-byte aligned tensors and exact_coords = 0 only");
  BMV_REQUIRE((int64_t)p->Hs * llabs(p->feat_y_stride) + (int64_t)p->Ws * llabs(p->feat_x_stride) < (1ll << 31),
              BMV_ERR_UNSUPPORTED_SHAPE, "bmv_cost_volume_var_multi: source map too large for 32-bit tap offsets");
  BMV_REQUIRE(p->out_bf16 >= 0 && p->out_bf16 <= 2, BMV_ERR_INVALID_ARGUMENT, "bmv_cost_volume_var_multi: out_bf16 must be 0, 1 or 2");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->out_bf16 == 1) return launch_cost_volume_multi<__nv_bfloat16>(*mp, st);
  if (p->out_bf16 == 2) return launch_cost_volume_multi<__half>(*mp, st);
  return launch_cost_volume_multi<float>(*mp, st);
}

extern "C" BMV_API int bmv_volume_scale(const bmv_volume_scale_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_volume_scale");
  using namespace bmv;
  BMV_REQUIRE(p && p->x && p->scale, BMV_ERR_INVALID_ARGUMENT, "bmv_volume_scale: null pointer");
  const int vec = p->x_half ? 8 : 4;
  BMV_REQUIRE(p->n >= vec && p->n % vec == 0 && ((uintptr_t)p->x & 15) == 0, BMV_ERR_INVALID_ARGUMENT,
              "bmv_volume_scale: n must be a positive multiple of %d and x 16-byte aligned", vec);
  BMV_REQUIRE(p->target > 0.f && p->target <= 65504.f, BMV_ERR_INVALID_ARGUMENT, "bmv_volume_scale: target must be in (0, 65504]");
  const int64_t want = ceil_div64(p->n / vec, 256 * 4);
  const unsigned blocks = (unsigned)(want < 8 * kNumSMs ? want : 8 * kNumSMs);
  if (p->x_half) volume_scale_kernel<__half><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  else volume_scale_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_volume_scale");
}

extern "C" BMV_API int bmv_depth_planes_first(const bmv_depth_planes_first_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_depth_planes_first");
  using namespace bmv;
  BMV_REQUIRE(p && p->near_far && p->t && p->planes && p->near_far_out, BMV_ERR_INVALID_ARGUMENT,
              "bmv_depth_planes_first: null pointer");
  BMV_REQUIRE(p->D >= 1 && p->h >= 1 && p->w >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_depth_planes_first: bad size");
  int n = max(p->D, p->h * p->w);
  depth_planes_first_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_depth_planes_first");
}

extern "C" BMV_API int bmv_depth_planes_next(const bmv_depth_planes_next_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_depth_planes_next");
  using namespace bmv;
  BMV_REQUIRE(p && p->depth && p->std && p->near_far && p->t && p->planes && p->near_far_out,
              BMV_ERR_INVALID_ARGUMENT, "bmv_depth_planes_next: null pointer");
  BMV_REQUIRE(p->D >= 1 && p->h >= 1 && p->w >= 1 && p->h0 >= 1 && p->w0 >= 1, BMV_ERR_INVALID_ARGUMENT,
              "bmv_depth_planes_next: bad size");
  BMV_REQUIRE(p->batch >= 0 && p->batch <= 65535, BMV_ERR_INVALID_ARGUMENT, "bmv_depth_planes_next: bad batch");
  int n = p->h * p->w;
  depth_planes_next_kernel<<<dim3((n + 255) / 256, p->batch > 1 ? p->batch : 1), 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_depth_planes_next");
}
