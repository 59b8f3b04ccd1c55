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
// `mid` is computed in fp32 with exactly the arithmetic of fpn.cu; the 3x3 convolution runs on tensor
// cores (mma.sync.m16n8k16, fp16 operands, fp32 accumulation: TF32-class — the host routes here only
// when torch.backends.cudnn.allow_tf32 is set, see inference_plan.py).
#include "bmv_internal.cuh"
#include "conv_mma.cuh"

namespace bmv {

constexpr int kFfThreads = 256;
constexpr int kFfTY = 8, kFfTX = 64, kFfWY = 4;                          // output tile per CTA; rows per warp job
constexpr int kFfHY = kFfTY + 2, kFfHX = kFfTX + 2;
constexpr int kFfRowB = kFfHX * 64;                                      // 32 fp16 channels per staged pixel
constexpr int kFfTileBytes = kFfHY * kFfRowB;

__device__ __forceinline__ int ff_swz(int v) { return (v >> 1) & 3; }

// OUT16: the instantiation that can write the fp16 version of the output (p.out16; p.out optional).  It is a template
// switch, not a run-time test: the extra epilogue code cost the plain fp32 instantiation 16 us (321 -> 337 us, B200).
template <int CIN, int NT, bool BREG, bool OUT16>
__global__ void __launch_bounds__(kFfThreads, NT == 1 ? 3 : 2) fpn_topdown_smooth_kernel(bmv_fpn_fused_params p) {
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
  __syncthreads();
  const int tiles_x = (p.W + kFfTX - 1) / kFfTX;
  const int x0 = (blockIdx.x % tiles_x) * kFfTX, y0 = (blockIdx.x / tiles_x) * kFfTY, n = blockIdx.y;
  const int Hp = p.H / 2, Wp = p.W / 2;
  // ---- phase 1: mid tile (+1 halo; zero outside the image = the convolution's padding) -> fp16 in shared memory
  {
    const float* lat = p.lateral_in + (int64_t)n * p.H * p.W * CIN;
    const float* prev = p.prev + (int64_t)n * Hp * Wp * 32;
    float* mid = p.mid ? p.mid + (int64_t)n * p.H * p.W * 32 : nullptr;
    const int cg = threadIdx.x & 7;
    const float4 bb = *reinterpret_cast<const float4*>(sW + 32 * CIN + cg * 4);
    // the lane's 4 x CIN lateral weights live in registers: as shared-memory operands (one LDS.128 per input
    // channel per item, 4 wavefronts each) they were 47 % of the L1 data-pipe wavefronts of this kernel (ncu)
    float4 wreg[CIN];
#pragma unroll
    for (int i = 0; i < CIN; ++i) wreg[i] = *reinterpret_cast<const float4*>(sW + i * 32 + cg * 4);
    const float sy = up_scale(Hp, p.H), sx = up_scale(Wp, p.W);
    for (int pi = threadIdx.x >> 3; pi < kFfHY * kFfHX; pi += kFfThreads / 8) {
      const int hy = pi / kFfHX, hx = pi - hy * kFfHX;
      const int y = y0 + hy - 1, x = x0 + hx - 1;
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
        const float* in = lat + ((int64_t)y * p.W + x) * CIN;
        float v[CIN];
#pragma unroll
        for (int i = 0; i < CIN; i += 4) {
          const float4 tt = __ldg(reinterpret_cast<const float4*>(in + i));
          v[i] = tt.x; v[i + 1] = tt.y; v[i + 2] = tt.z; v[i + 3] = tt.w;
        }
        const UpCoord uy = up_coord_scaled(y, Hp, sy), ux = up_coord_scaled(x, Wp, sx);
        const float* pb = prev + cg * 4;
        const float4 a = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i0 * Wp + ux.i0) * 32));
        const float4 b = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i0 * Wp + ux.i1) * 32));
        const float4 c = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i1 * Wp + ux.i0) * 32));
        const float4 d = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i1 * Wp + ux.i1) * 32));
        float acc[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int i = 0; i < CIN; ++i) {
          const float4 w = wreg[i];
          acc[0] = fmaf(w.x, v[i], acc[0]); acc[1] = fmaf(w.y, v[i], acc[1]);
          acc[2] = fmaf(w.z, v[i], acc[2]); acc[3] = fmaf(w.w, v[i], acc[3]);
        }
        r.x = (uy.l0 * (ux.l0 * a.x + ux.l1 * b.x) + uy.l1 * (ux.l0 * c.x + ux.l1 * d.x)) + acc[0];
        r.y = (uy.l0 * (ux.l0 * a.y + ux.l1 * b.y) + uy.l1 * (ux.l0 * c.y + ux.l1 * d.y)) + acc[1];
        r.z = (uy.l0 * (ux.l0 * a.z + ux.l1 * b.z) + uy.l1 * (ux.l0 * c.z + ux.l1 * d.z)) + acc[2];
        r.w = (uy.l0 * (ux.l0 * a.w + ux.l1 * b.w) + uy.l1 * (ux.l0 * c.w + ux.l1 * d.w)) + acc[3];
        if (mid && hy >= 1 && hy <= kFfTY && hx >= 1 && hx <= kFfTX)
          *reinterpret_cast<float4*>(mid + ((int64_t)y * p.W + x) * 32 + cg * 4) = r;
      }
      *reinterpret_cast<uint2*>(tile + hy * kFfRowB + hx * 64 + (((cg >> 1) ^ ff_swz(hx)) << 4) + (cg & 1) * 8) = pack_half4(r);
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

template <int CIN, int NT, bool BREG, bool OUT16>
static int launch_ff_t(const bmv_fpn_fused_params& p, cudaStream_t st) {
  const size_t smem = (size_t)kFfTileBytes + (size_t)3 * 6 * NT * 32 * 8 + (size_t)(32 * CIN + 32) * 4;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(fpn_topdown_smooth_kernel<CIN, NT, BREG, OUT16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      // 68 % of the 228 KB: room for the resident CTAs' tiles, the rest stays L1 for the prev / lateral taps
      // (measured on B200: 329 us with the maximum carve-out, 315 us with 64-72 %, 404 us at 50 %)
      e = cudaFuncSetAttribute(fpn_topdown_smooth_kernel<CIN, NT, BREG, OUT16>, cudaFuncAttributePreferredSharedMemoryCarveout, 68);
    if (e != cudaSuccess) {
      set_error("bmv_fpn_topdown_smooth: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const dim3 grid((unsigned)(((p.W + kFfTX - 1) / kFfTX) * ((p.H + kFfTY - 1) / kFfTY)), (unsigned)p.N);
  fpn_topdown_smooth_kernel<CIN, NT, BREG, OUT16><<<grid, kFfThreads, smem, st>>>(p);
  return check_launch("bmv_fpn_topdown_smooth");
}

template <int CIN, int NT, bool BREG>
static int launch_ff(const bmv_fpn_fused_params& p, cudaStream_t st) {
  return p.out16 ? launch_ff_t<CIN, NT, BREG, true>(p, st) : launch_ff_t<CIN, NT, BREG, false>(p, st);
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
  BMV_REQUIRE(((uintptr_t)p->prev & 15) == 0 && ((uintptr_t)p->lateral_in & 15) == 0 && ((uintptr_t)p->out & 7) == 0 &&
                  ((uintptr_t)p->wfrag & 15) == 0 && ((uintptr_t)p->mid & 15) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_topdown_smooth: tensors must be 16-byte aligned");
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
