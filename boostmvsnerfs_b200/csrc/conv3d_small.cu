// Tensor-core 3x3x3 convolutions for the LOW-RESOLUTION core of the kept 3-D cost regularisers: conv3 (16 -> 32, stride 2),
// conv4 (32 -> 32), conv5 (32 -> 64, stride 2), conv6 (64 -> 64) and the transposed conv7 (64 -> 32, stride 2, + skip)
// (reference lib/networks/enerf/cost_reg_net.py:14-24,58-75; BN folded by the caller).  These layers work on volumes of
// 8 k .. 65 k voxels (<= 4 MB of activations): on cuDNN each costs 8-22 us for < 2 us of tensor work — seven library
// launches, 81 us, in cost_reg_1 and two, 31 us, in cost_reg_0 (profiles/round2v_launches.csv).
//
// Direct implicit GEMM without a staged tile — the activations of a whole layer fit the L1s, so shared-memory staging
// would only add a barrier to kernels that run for a few microseconds: a warp owns 16 consecutive output voxels along x
// and NTW n-tiles (8 output channels each) and walks the 27 taps x Cin / 16 k-steps; every lane loads its A-fragment
// entries (two fp16 channels = 4 bytes) straight from global memory with the convolution's padding as load predicates,
// and the host-arranged B fragments with one 8-byte load per MMA.  Stride 2 addresses every second voxel.  The
// TRANSPOSED layer (k3, s2, p1, op1: out[o] = sum_k x[(o + 1 - k) / 2] w[k] over the k with o + 1 - k even) is the same
// loop over the 1 .. 8 taps that reach the warp's output-parity class; a warp's 16 voxels share their x parity.
// Same precision contract as conv3d_mma.cu (fp16 operands, fp32 accumulation: TF32-class gating by the host).
#include <cuda_fp16.h>

#include "bmv_internal.cuh"
#include "conv_mma.cuh"

namespace bmv {

constexpr int kCsThreads = 128;

template <int CIN, int NTW, bool TRANS>
__global__ void __launch_bounds__(kCsThreads) conv3d_small_kernel(bmv_conv3d_small_params p, int Do, int Ho, int Wo, int xtiles, int cgroups) {
  constexpr int KS = CIN / 16;
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // ---- which (n, od, oy, x tile [, x parity], channel group)
  int64_t job = (int64_t)blockIdx.x * (kCsThreads / 32) + (threadIdx.x >> 5);
  const int cg = (int)(job % cgroups); job /= cgroups;
  const int xt = (int)(job % xtiles); job /= xtiles;
  const int oy = (int)(job % Ho); job /= Ho;
  const int od = (int)(job % Do); job /= Do;
  const int n = (int)job;
  if (n >= p.N) return;                                                  // warp-uniform
  const int S = TRANS ? 1 : p.stride;
  // output x of fragment rows g (e = 0) and g + 8 (e = 1); transposed: the tile covers ONE x parity
  int px = 0, m0 = xt * 16;
  if (TRANS) { const int half_tiles = xtiles >> 1; px = xt >= half_tiles; m0 = (xt - px * half_tiles) * 16; }
  int ox[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) ox[e] = TRANS ? 2 * (m0 + g + 8 * e) + px : m0 + g + 8 * e;
  float acc[NTW][4];
#pragma unroll
  for (int nt = 0; nt < NTW; ++nt) {
    const int c = (cg * NTW + nt) * 8 + 2 * t;
    const float b0 = p.bias ? __ldg(p.bias + c) : 0.f, b1 = p.bias ? __ldg(p.bias + c + 1) : 0.f;
    acc[nt][0] = b0; acc[nt][1] = b1; acc[nt][2] = b0; acc[nt][3] = b1;
  }
  const __half* xin = reinterpret_cast<const __half*>(p.x) + (int64_t)n * p.x_n_stride + 2 * t;
  const uint2* wfrag = reinterpret_cast<const uint2*>(p.wfrag);
  const int NT = p.Cout / 8;
#pragma unroll 1
  for (int kd = 0; kd < 3; ++kd) {
    int id;
    if (TRANS) { if ((od + 1 - kd) & 1) continue; id = (od + 1 - kd) >> 1; } else id = od * S - 1 + kd;
    if (id < 0 || id >= p.D) continue;                                   // warp-uniform
#pragma unroll 1
    for (int ky = 0; ky < 3; ++ky) {
      int iy;
      if (TRANS) { if ((oy + 1 - ky) & 1) continue; iy = (oy + 1 - ky) >> 1; } else iy = oy * S - 1 + ky;
      if (iy < 0 || iy >= p.H) continue;
      const __half* xrow = xin + (int64_t)id * p.x_d_stride + (int64_t)iy * p.x_y_stride;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        if (TRANS && ((px + 1 - kx) & 1)) continue;                      // warp-uniform (one x parity per warp)
        const __half* src[2];
        bool ok[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int ix = TRANS ? (ox[e] + 1 - kx) >> 1 : ox[e] * S - 1 + kx;
          ok[e] = ix >= 0 && ix < p.W && ox[e] < Wo;
          src[e] = xrow + (int64_t)(ok[e] ? ix : 0) * p.x_x_stride;
        }
        const int tap = (kd * 3 + ky) * 3 + kx;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t a[4];
          a[0] = ok[0] ? __ldg(reinterpret_cast<const uint32_t*>(src[0] + ks * 16)) : 0u;
          a[1] = ok[1] ? __ldg(reinterpret_cast<const uint32_t*>(src[1] + ks * 16)) : 0u;
          a[2] = ok[0] ? __ldg(reinterpret_cast<const uint32_t*>(src[0] + ks * 16 + 8)) : 0u;
          a[3] = ok[1] ? __ldg(reinterpret_cast<const uint32_t*>(src[1] + ks * 16 + 8)) : 0u;
          const uint2* wb = wfrag + ((int64_t)(tap * KS + ks) * NT + cg * NTW) * 32 + lane;
#pragma unroll
          for (int nt = 0; nt < NTW; ++nt) {
            const uint2 bw = __ldg(wb + nt * 32);
            hmma16816(acc[nt], a, bw.x, bw.y);
          }
        }
      }
    }
  }
  // ---- epilogue: (+ skip) (ReLU) store
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    if (ox[e] >= Wo) continue;
    const int64_t vox = (((int64_t)n * Do + od) * Ho + oy) * Wo + ox[e];
#pragma unroll
    for (int nt = 0; nt < NTW; ++nt) {
      const int c = (cg * NTW + nt) * 8 + 2 * t;
      float v0 = acc[nt][2 * e], v1 = acc[nt][2 * e + 1];
      if (p.skip) {
        const float2 s = __half22float2(*reinterpret_cast<const __half2*>(reinterpret_cast<const __half*>(p.skip) + vox * p.Cout + c));
        v0 += s.x; v1 += s.y;
      }
      if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); }
      if (p.out_half) *reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(p.out) + vox * p.Cout + c) = pack_half2_sat(v0, v1);
      else *reinterpret_cast<float2*>(reinterpret_cast<float*>(p.out) + vox * p.Cout + c) = make_float2(v0, v1);
    }
  }
}

template <int CIN, int NTW, bool TRANS>
static int launch_cs(const bmv_conv3d_small_params& p, cudaStream_t st) {
  const int S = p.stride;
  const int Do = TRANS ? 2 * p.D : (p.D - 1) / S + 1, Ho = TRANS ? 2 * p.H : (p.H - 1) / S + 1, Wo = TRANS ? 2 * p.W : (p.W - 1) / S + 1;
  const int xtiles = TRANS ? 2 * ((p.W + 15) / 16) : (Wo + 15) / 16;
  const int cgroups = p.Cout / (8 * NTW);
  const int64_t jobs = (int64_t)p.N * Do * Ho * xtiles * cgroups;
  const int64_t blocks = ceil_div64(jobs, kCsThreads / 32);
  if (blocks > 0x7fffffffll) {
    set_error("bmv_conv3d_small: volume too large");
    return BMV_ERR_UNSUPPORTED_SHAPE;
  }
  conv3d_small_kernel<CIN, NTW, TRANS><<<(unsigned)blocks, kCsThreads, 0, st>>>(p, Do, Ho, Wo, xtiles, cgroups);
  return check_launch("bmv_conv3d_small");
}

}  // namespace bmv

extern "C" BMV_API int bmv_conv3d_small(const bmv_conv3d_small_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_conv3d_small");
  using namespace bmv;
  BMV_REQUIRE(p && p->x && p->wfrag && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_small: null pointer");
  BMV_REQUIRE(p->N >= 1 && p->D >= 1 && p->H >= 1 && p->W >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_small: bad size");
  BMV_REQUIRE(p->x_n_stride % 2 == 0 && p->x_d_stride % 2 == 0 && p->x_y_stride % 2 == 0 && p->x_x_stride % 2 == 0 && ((uintptr_t)p->x & 3) == 0 &&
                  ((uintptr_t)p->wfrag & 7) == 0 && ((uintptr_t)p->out & 7) == 0 && ((uintptr_t)p->skip & 3) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_small: fp16 channels-last input with 4-byte aligned voxels, 8-byte aligned wfrag / out");
  BMV_REQUIRE(p->transposed ? p->stride == 2 : (p->stride == 1 || p->stride == 2), BMV_ERR_INVALID_ARGUMENT,
              "bmv_conv3d_small: stride 1 or 2 (transposed: 2)");
  cudaStream_t st = (cudaStream_t)stream;
  if (!p->transposed && p->Cin == 16 && p->Cout == 32) return launch_cs<16, 4, false>(*p, st);
  if (!p->transposed && p->Cin == 32 && p->Cout == 32) return launch_cs<32, 4, false>(*p, st);
  if (!p->transposed && p->Cin == 32 && p->Cout == 64) return launch_cs<32, 4, false>(*p, st);
  if (!p->transposed && p->Cin == 64 && p->Cout == 64) return launch_cs<64, 4, false>(*p, st);
  if (p->transposed && p->Cin == 64 && p->Cout == 32) return launch_cs<64, 4, true>(*p, st);
  set_error("bmv_conv3d_small: (Cin=%d, Cout=%d, transposed=%d) not instantiated (16->32, 32->32, 32->64, 64->64; transposed 64->32)",
            p->Cin, p->Cout, p->transposed);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}

// words (uint32) of the fragment-ordered weight buffer ([tap 27][k-step][n-tile][lane][2]), -1 if not instantiated
extern "C" BMV_API int bmv_conv3d_small_weight_words(int Cin, int Cout, int transposed) {
  const bool ok = transposed ? (Cin == 64 && Cout == 32)
                             : ((Cin == 16 && Cout == 32) || (Cin == 32 && Cout == 32) || (Cin == 32 && Cout == 64) || (Cin == 64 && Cout == 64));
  return ok ? 27 * (Cin / 16) * (Cout / 8) * 32 * 2 : -1;
}
