// Fused top-down step + smoothing convolution of the 2-D feature pyramid:
//     mid = bilinear_up2x(prev, align_corners=True) + conv1x1(lateral_in) + b_lat        (32 channels)
//     out = conv3x3(mid, pad 1) + b_smooth                                                (8 or 16 channels)
// (`_upsample_add(x, lat(c))` followed by `smooth(x)`, reference lib/networks/enerf/feature_net.py:24-47.)
//
// Why: the 32-channel full-resolution `mid` is 401 MB for the 6 source views of C2.  As separate
// launches it is written once and read once (0.81 + 0.18 ms measured on B200); fused it only ever
// exists as an fp16 tile in shared memory and the step is bounded by reading prev / lateral_in and
// writing the 8-channel output.  At half resolution `mid` is also the next step's `prev`, so it is
// additionally written out (exact fp32) when p.mid is given.
//
// `mid` is computed to fp32 accuracy (upsample: fp32 with product weights, the value of fpn.cu up to the last ulps; 1x1
// lateral: tensor cores with both operands split into fp16 hi + lo, ~1e-6 relative); the 3x3 convolution runs on tensor cores with plain fp16 operands
// (mma.sync.m16n8k16, fp32 accumulation: TF32-class — the host routes here only when
// torch.backends.cudnn.allow_tf32 is set, see inference_plan.py).
#include <stdlib.h>

#include "bmv_internal.cuh"
#include "conv_mma.cuh"

namespace bmv {

constexpr int kFfThreads = 256;
constexpr int kFfTY = 8, kFfTX = 64, kFfWY = 4;                          // output tile per CTA; rows per warp job
constexpr int kFfHY = kFfTY + 2, kFfHX = kFfTX + 2;
constexpr int kFfRowB = kFfHX * 64;                                      // 32 fp16 channels per staged pixel
constexpr int kFfTileBytes = kFfHY * kFfRowB;

__device__ __forceinline__ int ff_swz(int v) { return (v >> 1) & 3; }

// 8 consecutive fp32 channels with one 256-bit load (32-byte aligned)
__device__ __forceinline__ void ldg8_f32(const float* q, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(q));
}

// OUT16: the instantiation that can write the fp16 version of the output (p.out16; p.out optional).  It is a template
// switch, not a run-time test: the extra epilogue code cost the plain fp32 instantiation 16 us (321 -> 337 us, B200).
template <int CIN, int NT, bool BREG, bool OUT16, int MINB, bool LAT16>
__global__ void __launch_bounds__(kFfThreads, MINB) fpn_topdown_smooth_kernel(bmv_fpn_fused_params p) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* tile = smem;
  uint2* wfrag = reinterpret_cast<uint2*>(smem + kFfTileBytes);
  constexpr int W_WORDS = 3 * 6 * NT * 32 * 2;
  float* sW = reinterpret_cast<float*>(smem + kFfTileBytes + W_WORDS * 4);          // [CIN][32] + bias[32]
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wfrag);
    uint4* dst = reinterpret_cast<uint4*>(wfrag);
    for (int i = threadIdx.x; i < W_WORDS / 4; i += kFfThreads) dst[i] = __ldg(src + i);
    for (int i = threadIdx.x; i < 32 * CIN; i += kFfThreads) sW[(i % CIN) * 32 + i / CIN] = __ldg(p.lat_weight + i);
    if (threadIdx.x < 32) sW[32 * CIN + threadIdx.x] = p.lat_bias ? __ldg(p.lat_bias + threadIdx.x) : 0.f;
  }
  const int tiles_x = (p.W + kFfTX - 1) / kFfTX;
  const int x0 = (blockIdx.x % tiles_x) * kFfTX, y0 = (blockIdx.x / tiles_x) * kFfTY, n = blockIdx.y;
  const int Hp = p.H / 2, Wp = p.W / 2;
  // Per-CTA coordinate tables of the 10 halo rows / 66 halo columns (clamped into the image): element offsets of the two
  // `prev` rows / columns a pixel interpolates between, the two weights, and the offset of its lateral-input row / column.
  // ncu of the version that derived all this per pixel: 186 of the ~380 instructions of a 16-pixel segment were index
  // arithmetic (64-bit address chains, clamps, int <-> float conversions of the align_corners coordinates).
  __shared__ int4 s_row[kFfHY], s_col[kFfHX];
  __shared__ int s_rowlat[kFfHY], s_collat[kFfHX];
  if (threadIdx.x < kFfHY) {
    const int cy = min(max(y0 + (int)threadIdx.x - 1, 0), p.H - 1);
    const UpCoord u = up_coord_scaled(cy, Hp, up_scale(Hp, p.H));
    s_row[threadIdx.x] = make_int4(u.i0 * Wp * 32, u.i1 * Wp * 32, __float_as_int(u.l0), __float_as_int(u.l1));
    s_rowlat[threadIdx.x] = cy * p.W * CIN;
  } else if (threadIdx.x >= 32 && threadIdx.x < 32 + kFfHX) {
    const int hx = threadIdx.x - 32;
    const int cx = min(max(x0 + hx - 1, 0), p.W - 1);
    const UpCoord u = up_coord_scaled(cx, Wp, up_scale(Wp, p.W));
    s_col[hx] = make_int4(u.i0 * 32, u.i1 * 32, __float_as_int(u.l0), __float_as_int(u.l1));
    s_collat[hx] = cx * CIN;
  }
  __syncthreads();
  // ---- phase 1: mid tile (+1 halo; zero outside the image = the convolution's padding) -> fp16 in shared memory.
  // The 1x1 lateral convolution runs on the tensor cores too (round 2; the CUDA-core version spent ~1.2 k thread
  // instructions per pixel here, 8 lanes x 150, and made the kernel issue-bound at 62 %): a warp takes 16 consecutive
  // halo pixels as the M dimension of mma.sync.m16n8k16, 4 n-tiles = the 32 channels.  fp32 accuracy is kept by
  // splitting both operands into fp16 hi + lo (lo pre-multiplied by 2^11 so that it stays a normal number):
  //     lat = [x_hi * W_hi] + 2^-11 * [x_lo' * W_hi + x_hi * W_lo']     (the dropped lo * lo term is ~2^-22 relative)
  // in two fp32 accumulators.  The columns of n-tile nt are PERMUTED (column n <-> channel (n / 2) * 8 + nt * 2 + n % 2)
  // so that a lane's C fragments hold 8 CONSECUTIVE channels 8t .. 8t+7 of its two pixels: every `prev` tap is one
  // 256-bit load, the fp16 tile entry one 16-byte store.
  {
    const float* lat = p.lateral_in + (LAT16 ? 0 : (int64_t)n * p.H * p.W * CIN);
    const __half* lat16 = reinterpret_cast<const __half*>(p.lateral_in) + (int64_t)n * p.H * p.W * CIN;
    const float* prev = p.prev + (int64_t)n * Hp * Wp * 32;
    float* mid = p.mid ? p.mid + (int64_t)n * p.H * p.W * 32 : nullptr;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    constexpr int KK = CIN / 8;                                         // 8-channel groups of the lateral input
    constexpr float kLoUp = 2048.f, kLoDown = 1.f / 2048.f;
    uint32_t whi[4][KK], wlo[4][KK];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int ch = (g >> 1) * 8 + nt * 2 + (g & 1);                   // channel of B-fragment column g
#pragma unroll
      for (int kk = 0; kk < KK; ++kk) {
        const float w0 = sW[(2 * t + 8 * kk) * 32 + ch], w1 = sW[(2 * t + 8 * kk + 1) * 32 + ch];
        const __half h0 = half_sat(w0), h1 = half_sat(w1);
        whi[nt][kk] = pack_half2_sat(w0, w1);
        wlo[nt][kk] = pack_half2_sat((w0 - __half2float(h0)) * kLoUp, (w1 - __half2float(h1)) * kLoUp);
      }
    }
    constexpr int NPIX = kFfHY * kFfHX, NSEG = (NPIX + 15) / 16;
    const float* prevb = prev + t * 8;
    for (int seg = warp; seg < NSEG; seg += kFfThreads / 32) {
      // Every global load of the segment is issued up front with CLAMPED (always valid) addresses — the lateral inputs of
      // both pixels and their 2 x 4 `prev` taps; out-of-image pixels are zeroed when the tile entry is built.
      int hyv[2], hxv[2], poff[2];
      bool ok[2];
      float w00[2], w01[2], w10[2], w11[2];
      uint32_t ahi[2][KK], alo[2][KK];
      float2 lv[2][KK];
      uint32_t lh[2][KK];
      float4 tap[2][4][2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int pi = min(seg * 16 + g + 8 * e, NPIX - 1);
        hyv[e] = pi / kFfHX; hxv[e] = pi - hyv[e] * kFfHX;
        const int py = y0 + hyv[e] - 1, px = x0 + hxv[e] - 1;
        ok[e] = (unsigned)py < (unsigned)p.H && (unsigned)px < (unsigned)p.W;
        poff[e] = (py * p.W + px) * 32 + t * 8;                         // `mid` element (used for in-image pixels only)
        const int4 r = s_row[hyv[e]], c = s_col[hxv[e]];
        const int loff = s_rowlat[hyv[e]] + s_collat[hxv[e]] + 2 * t;
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          if (LAT16) lh[e][kk] = __ldg(reinterpret_cast<const uint32_t*>(lat16 + loff + 8 * kk));
          else lv[e][kk] = __ldg(reinterpret_cast<const float2*>(lat + loff + 8 * kk));
        }
        ldg8_f32(prevb + (r.x + c.x), tap[e][0][0], tap[e][0][1]);
        ldg8_f32(prevb + (r.x + c.y), tap[e][1][0], tap[e][1][1]);
        ldg8_f32(prevb + (r.y + c.x), tap[e][2][0], tap[e][2][1]);
        ldg8_f32(prevb + (r.y + c.y), tap[e][3][0], tap[e][3][1]);
        const float yl0 = __int_as_float(r.z), yl1 = __int_as_float(r.w), xl0 = __int_as_float(c.z), xl1 = __int_as_float(c.w);
        w00[e] = yl0 * xl0; w01[e] = yl0 * xl1; w10[e] = yl1 * xl0; w11[e] = yl1 * xl1;
      }
#pragma unroll
      for (int e = 0; e < 2; ++e)
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          if (LAT16) {          // fp16 lateral input (what the stem / bmv_conv2d_k3 emit): exact operands, no lo part
            ahi[e][kk] = lh[e][kk];
            alo[e][kk] = 0u;
          } else {
            const float2 vv = lv[e][kk];
            const __half h0 = half_sat(vv.x), h1 = half_sat(vv.y);
            ahi[e][kk] = pack_half2_sat(vv.x, vv.y);
            alo[e][kk] = pack_half2_sat((vv.x - __half2float(h0)) * kLoUp, (vv.y - __half2float(h1)) * kLoUp);
          }
        }
      float chi[4][4], clo[4][4];
      const float4 lb0 = *reinterpret_cast<const float4*>(sW + 32 * CIN + t * 8), lb1 = *reinterpret_cast<const float4*>(sW + 32 * CIN + t * 8 + 4);
      const float lbias[4][2] = {{lb0.x, lb0.y}, {lb0.z, lb0.w}, {lb1.x, lb1.y}, {lb1.z, lb1.w}};
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        chi[nt][0] = lbias[nt][0]; chi[nt][1] = lbias[nt][1]; chi[nt][2] = lbias[nt][0]; chi[nt][3] = lbias[nt][1];
        clo[nt][0] = 0.f; clo[nt][1] = 0.f; clo[nt][2] = 0.f; clo[nt][3] = 0.f;
        if (KK == 1) {        // K = 16 holds [x_hi (8) | x_lo' (8)]: B = [W_hi ; 0] and [W_lo' ; W_hi]
          const uint32_t a[4] = {ahi[0][0], ahi[1][0], alo[0][0], alo[1][0]};
          hmma16816(chi[nt], a, whi[nt][0], 0u);
          hmma16816(clo[nt], a, wlo[nt][0], whi[nt][0]);
        } else {              // K = 16 channels: x_hi * W_hi, x_lo' * W_hi, x_hi * W_lo'
          const uint32_t a_h[4] = {ahi[0][0], ahi[1][0], ahi[0][KK - 1], ahi[1][KK - 1]};
          const uint32_t a_l[4] = {alo[0][0], alo[1][0], alo[0][KK - 1], alo[1][KK - 1]};
          hmma16816(chi[nt], a_h, whi[nt][0], whi[nt][KK - 1]);
          if (!LAT16) hmma16816(clo[nt], a_l, whi[nt][0], whi[nt][KK - 1]);
          hmma16816(clo[nt], a_h, wlo[nt][0], wlo[nt][KK - 1]);
        }
      }
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        if (seg * 16 + g + 8 * e >= NPIX) continue;
        float4 r0 = make_float4(0.f, 0.f, 0.f, 0.f), r1 = r0;
        if (ok[e]) {
          // bilinear blend with the four PRODUCT weights (5 FMAs per channel incl. the lateral term; the nested form of
          // fpn.cu takes 8): same value up to the last ulps
          const int o = 2 * e;
#define BMV_FF_BLEND(dst, f, q, nt, j)                                                                                         \
          dst = fmaf(w00[e], tap[e][0][q].f, fmaf(w01[e], tap[e][1][q].f, fmaf(w10[e], tap[e][2][q].f,                       \
                     fmaf(w11[e], tap[e][3][q].f, fmaf(clo[nt][o + j], kLoDown, chi[nt][o + j])))))
          BMV_FF_BLEND(r0.x, x, 0, 0, 0); BMV_FF_BLEND(r0.y, y, 0, 0, 1); BMV_FF_BLEND(r0.z, z, 0, 1, 0); BMV_FF_BLEND(r0.w, w, 0, 1, 1);
          BMV_FF_BLEND(r1.x, x, 1, 2, 0); BMV_FF_BLEND(r1.y, y, 1, 2, 1); BMV_FF_BLEND(r1.z, z, 1, 3, 0); BMV_FF_BLEND(r1.w, w, 1, 3, 1);
#undef BMV_FF_BLEND
          if (mid && hyv[e] >= 1 && hyv[e] <= kFfTY && hxv[e] >= 1 && hxv[e] <= kFfTX) {
            *reinterpret_cast<float4*>(mid + poff[e]) = r0;
            *reinterpret_cast<float4*>(mid + poff[e] + 4) = r1;
          }
        }
        uint4 pk;
        pk.x = pack_half2_sat(r0.x, r0.y); pk.y = pack_half2_sat(r0.z, r0.w);
        pk.z = pack_half2_sat(r1.x, r1.y); pk.w = pack_half2_sat(r1.z, r1.w);
        *reinterpret_cast<uint4*>(tile + hyv[e] * kFfRowB + hxv[e] * 64 + ((t ^ ff_swz(hxv[e])) << 4)) = pk;
      }
    }
  }
  __syncthreads();
  // ---- phase 2: 3x3 convolution 32 -> 8*NT on tensor cores; a warp owns kFfWY rows x 16 pixels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;
  uint2 breg[BREG ? 18 * NT : 1];
  if (BREG) {
#pragma unroll
    for (int i = 0; i < 18 * NT; ++i) breg[i] = wfrag[i * 32 + lane];
  }
  constexpr int JOBS = (kFfTY / kFfWY) * (kFfTX / 16);
  for (int job = warp; job < JOBS; job += kFfThreads / 32) {
    const int mx = job % (kFfTX / 16), yb = (job / (kFfTX / 16)) * kFfWY;
    if (y0 + yb >= p.H || x0 + mx * 16 >= p.W) continue;
    float acc[kFfWY][NT][4];
#pragma unroll
    for (int oy = 0; oy < kFfWY; ++oy)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + 2 * t;
        const float b0 = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f, b1 = (p.bias && c + 1 < p.Cout) ? __ldg(p.bias + c + 1) : 0.f;
        acc[oy][nt][0] = b0; acc[oy][nt][1] = b1; acc[oy][nt][2] = b0; acc[oy][nt][3] = b1;
      }
    uint32_t aoff[6];
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      const int v = mx * 16 + lrow + (j >> 1);
      aoff[j] = tile_s + v * 64 + ((((j & 1) * 2 + lhi) ^ ff_swz(v)) << 4);
    }
#pragma unroll
    for (int py = 0; py < kFfWY + 2; ++py) {
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        uint32_t a[4];
        ldmatrix_x4(a, aoff[j] + (yb + py) * kFfRowB);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int oy = py - dy;
          if (oy < 0 || oy >= kFfWY) continue;
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
            const int wi = (dy * 6 + j) * NT + nt;
            const uint2 bw = BREG ? breg[BREG ? wi : 0] : wfrag[wi * 32 + lane];
            hmma16816(acc[oy][nt], a, bw.x, bw.y);
          }
        }
      }
    }
    float* out = p.out + (int64_t)n * p.H * p.W * p.Cout;
    const int gx0 = x0 + mx * 16 + g, gx1 = gx0 + 8;
#pragma unroll
    for (int oy = 0; oy < kFfWY; ++oy) {
      const int gy = y0 + yb + oy;
      if (gy >= p.H) continue;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + 2 * t;
        if (c + 1 < p.Cout + 1 && c < p.Cout) {                         // Cout is even (8 or 16): whole pairs
          if (!OUT16 || p.out) {
            if (gx0 < p.W) *reinterpret_cast<float2*>(out + ((int64_t)gy * p.W + gx0) * p.Cout + c) = make_float2(acc[oy][nt][0], acc[oy][nt][1]);
            if (gx1 < p.W) *reinterpret_cast<float2*>(out + ((int64_t)gy * p.W + gx1) * p.Cout + c) = make_float2(acc[oy][nt][2], acc[oy][nt][3]);
          }
          if (OUT16 && p.out16) {
            __half* o16 = reinterpret_cast<__half*>(p.out16) + (int64_t)n * p.H * p.W * p.Cout;
            if (gx0 < p.W) *reinterpret_cast<uint32_t*>(o16 + ((int64_t)gy * p.W + gx0) * p.Cout + c) = pack_half2_sat(acc[oy][nt][0], acc[oy][nt][1]);
            if (gx1 < p.W) *reinterpret_cast<uint32_t*>(o16 + ((int64_t)gy * p.W + gx1) * p.Cout + c) = pack_half2_sat(acc[oy][nt][2], acc[oy][nt][3]);
          }
        }
      }
    }
  }
}

template <int CIN, int NT, bool BREG, bool OUT16, int MINB, bool LAT16>
static int launch_ff_l(const bmv_fpn_fused_params& p, cudaStream_t st) {
  const size_t smem = (size_t)kFfTileBytes + (size_t)3 * 6 * NT * 32 * 8 + (size_t)(32 * CIN + 32) * 4;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(fpn_topdown_smooth_kernel<CIN, NT, BREG, OUT16, MINB, LAT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      // 68 % of the 228 KB: room for the resident CTAs' tiles, the rest stays L1 for the prev / lateral taps
      // (measured on B200: 329 us with the maximum carve-out, 315 us with 64-72 %, 404 us at 50 %)
      e = cudaFuncSetAttribute(fpn_topdown_smooth_kernel<CIN, NT, BREG, OUT16, MINB, LAT16>, cudaFuncAttributePreferredSharedMemoryCarveout, 68);
    if (e != cudaSuccess) {
      set_error("bmv_fpn_topdown_smooth: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const dim3 grid((unsigned)(((p.W + kFfTX - 1) / kFfTX) * ((p.H + kFfTY - 1) / kFfTY)), (unsigned)p.N);
  fpn_topdown_smooth_kernel<CIN, NT, BREG, OUT16, MINB, LAT16><<<grid, kFfThreads, smem, st>>>(p);
  return check_launch("bmv_fpn_topdown_smooth");
}

template <int CIN, int NT, bool BREG, bool OUT16, int MINB>
static int launch_ff_t(const bmv_fpn_fused_params& p, cudaStream_t st) {
  return p.lat_half ? launch_ff_l<CIN, NT, BREG, OUT16, MINB, true>(p, st) : launch_ff_l<CIN, NT, BREG, OUT16, MINB, false>(p, st);
}

template <int CIN, int NT, bool BREG>
static int launch_ff(const bmv_fpn_fused_params& p, cudaStream_t st) {
  // resident CTAs per SM the kernel is compiled for: 3 (80 registers) or 2 (128); BMV_FF_MINB overrides (measurements)
  static const int minb_env = getenv("BMV_FF_MINB") ? atoi(getenv("BMV_FF_MINB")) : 0;
  const int minb = minb_env ? minb_env : 2;
  if constexpr (NT == 1) {
    if (minb == 3) return p.out16 ? launch_ff_t<CIN, NT, BREG, true, 3>(p, st) : launch_ff_t<CIN, NT, BREG, false, 3>(p, st);
  }
  return p.out16 ? launch_ff_t<CIN, NT, BREG, true, 2>(p, st) : launch_ff_t<CIN, NT, BREG, false, 2>(p, st);
}

}  // namespace bmv

extern "C" BMV_API int bmv_fpn_topdown_smooth(const bmv_fpn_fused_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_fpn_topdown_smooth");
  using namespace bmv;
  BMV_REQUIRE(p && p->prev && p->lateral_in && p->lat_weight && p->wfrag && (p->out || p->out16), BMV_ERR_INVALID_ARGUMENT,
              "bmv_fpn_topdown_smooth: null pointer");
  BMV_REQUIRE(((uintptr_t)p->out16 & 3) == 0, BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_topdown_smooth: out16 must be 4-byte aligned");
  BMV_REQUIRE(p->N >= 1 && p->N <= 65535 && p->H >= 2 && p->W >= 2 && p->H % 2 == 0 && p->W % 2 == 0, BMV_ERR_INVALID_ARGUMENT,
              "bmv_fpn_topdown_smooth: H, W must be even and >= 2");
  BMV_REQUIRE((int64_t)p->H * p->W <= (1ll << 25), BMV_ERR_UNSUPPORTED_SHAPE, "bmv_fpn_topdown_smooth: image too large for 32-bit offsets");
  BMV_REQUIRE(((uintptr_t)p->prev & 31) == 0 && ((uintptr_t)p->lateral_in & 15) == 0 && ((uintptr_t)p->out & 7) == 0 &&
                  ((uintptr_t)p->wfrag & 15) == 0 && ((uintptr_t)p->mid & 15) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_topdown_smooth: tensors must be 16-byte aligned (prev: 32-byte)");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->Cin == 8 && p->Cout == 8) return launch_ff<8, 1, true>(*p, st);
  if (p->Cin == 16 && p->Cout == 16) return launch_ff<16, 2, false>(*p, st);
  if (p->Cin == 16 && p->Cout == 8) return launch_ff<16, 1, true>(*p, st);
  if (p->Cin == 8 && p->Cout == 16) return launch_ff<8, 2, false>(*p, st);
  set_error("bmv_fpn_topdown_smooth: (Cin=%d, Cout=%d) not instantiated (Cin 8|16, Cout 8|16)", p->Cin, p->Cout);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}

extern "C" BMV_API int bmv_fpn_topdown_smooth_weight_words(int Cout) {
  return (Cout == 8 || Cout == 16) ? 3 * 6 * (Cout / 8) * 32 * 2 : -1;
}
