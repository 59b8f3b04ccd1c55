// Tensor-core 3x3 / pad-1 convolution (+bias, +ReLU) for the MIDDLE layers of the kept 2-D feature pyramid
// (conv1.x, conv2.x and the 1x1 top layer; reference lib/networks/enerf/feature_net.py:10-21,29-31), channels-last.
//
// Why (north_star: "unless ncu shows a hand-written tensor-core tile pays off"): with everything else optimised the
// seven library launches these layers took (four cuDNN convolutions, the top layer + its bias add, a space-to-depth
// copy) were 221 us of the 3.35 ms C2 frame (profiles/round2v_launches.csv), 30-45 us each for 20-25 us of memory
// traffic and < 10 us of tensor work.
//
// Same precision contract as conv3d_mma.cu: operands rounded to fp16, fp32 accumulation (TF32-class; the host routes
// here only when torch.backends.cudnn.allow_tf32 is set).
//
// Implicit GEMM on mma.sync.m16n8k16, M = 16 consecutive x pixels, N = 8 output channels per n-tile, one k-step = 16
// input channels of pixel x + dx.  A CTA stages an (16+2) x (32+2) input tile as fp16 in shared memory (zero outside the
// image = the padding; 16-byte chunks XOR-swizzled so that every ldmatrix is conflict-free) and its 8 warps each own
// 4 rows x 16 pixels x ALL output channels.  Loop order: k-step outermost — its 3 x NT weight fragments (the three dy)
// are loaded ONCE into registers and used by the six input rows a warp walks (a per-row order re-reads them 4 times and
// makes the load pipe, not the tensor pipe, the bound).
//   * Input modes: fp16 dense, or fp16 / fp32 SPACE-TO-DEPTH: the 5x5 / stride-2 layers are evaluated as 3x3 convolutions
//     over space-to-depth(2) of their input (inference_plan.S2DConv5x5: same products, regrouped) and the staging loop does
//     that regrouping — the (N, 4C, H/2, W/2) tensor is never materialised.  fp16 sources are staged with 16-byte cp.async
//     (zero-fill outside the image), fp32 sources through registers (converted on the way).
//   * Output channels are PERMUTED across the n-tiles (column n of n-tile nt <-> channel (n / 2) * 2NT + 2nt + n % 2, by
//     the host-side packing) so that a lane's C fragments hold 2NT CONSECUTIVE channels of a pixel: 16- / 32-byte stores.
//   * Optional fused 1x1 convolution (the FPN top layer on conv2.1's output): the ReLU'd C fragments are re-packed as A
//     fragments (registers only) and multiplied by the 32 x 32 matrix; the intermediate never reaches memory.
#include <cuda_fp16.h>
#include <stdlib.h>

#include "bmv_internal.cuh"
#include "conv_mma.cuh"

namespace bmv {

constexpr int kC2TX = 32, kC2WY = 4;
constexpr int kC2HX = kC2TX + 2;

// TY: output rows per CTA (16: 8 warps, 8: 4 warps — the quarter-resolution layers launch too few 16-row tiles to balance
// 148 SMs: 432 CTAs on 296 slots)
template <int CIN, int TY> struct C2Cfg {
  static constexpr int HY = TY + 2, THREADS = (TY / kC2WY) * 2 * 32;
  static constexpr int VS = 2 * CIN;                                     // bytes per staged pixel
  static constexpr int CH8 = CIN / 8;                                    // 16-byte chunks per pixel
  static constexpr int KPD = CIN / 16;                                   // k-steps per dx
  static constexpr int KS = 3 * KPD;
  static constexpr int ROWB = kC2HX * VS;
  static constexpr int TILE_BYTES = HY * ROWB;
  // chunk c of pixel v lives at chunk c ^ swz(v): the 8 rows of an ldmatrix 8x8 block hit 8 distinct bank groups
  __device__ static __forceinline__ int swz(int v) { return CIN == 16 ? ((v >> 2) & 1) : (CIN == 32 ? ((v >> 1) & 3) : (v & 7)); }
};

// MODE 1: fp16 dense input; MODE 2 / 3: fp32 / fp16 input read through space-to-depth(2)
template <int CIN, int NT, int MODE, bool FUSE, int TY>
__global__ void __launch_bounds__(C2Cfg<CIN, TY>::THREADS, TY == 16 ? 2 : 4) conv2d_k3_mma_kernel(bmv_conv2d_params p) {
  using Cfg = C2Cfg<CIN, TY>;
  constexpr int kC2Threads = Cfg::THREADS, kC2HY = Cfg::HY, kC2TY = TY;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* tile = smem;
  // weight fragments straight from global memory (<= 36 KB, L1-resident after the first tile of an SM): every k-step
  // reads its 3 x NT fragments once per warp, so a shared-memory copy would only cost its prologue and a resident CTA
  const uint2* wfrag = reinterpret_cast<const uint2*>(p.wfrag);
  const int tiles_x = (p.W + kC2TX - 1) / kC2TX;
  const int x0 = (blockIdx.x % tiles_x) * kC2TX, y0 = (blockIdx.x / tiles_x) * kC2TY, n = blockIdx.y;
  // ---- stage the input tile (+1 halo) as fp16
  if (MODE == 1 || MODE == 3) {
    // fp16 source: 16-byte cp.async per (pixel, chunk), zero-filled outside the image (src-size 0), nothing staged through
    // registers; all copies of a thread are in flight together
    constexpr int ITEMS = kC2HY * kC2HX * Cfg::CH8;
    const __half* xin = reinterpret_cast<const __half*>(p.x) + (int64_t)n * p.x_n_stride;
    const uint32_t tile_s0 = (uint32_t)__cvta_generic_to_shared(tile);
    for (int idx = threadIdx.x; idx < ITEMS; idx += kC2Threads) {
      const int pix = idx / Cfg::CH8, c8 = idx % Cfg::CH8;
      const int hy = pix / kC2HX, hx = pix - hy * kC2HX;
      const int y = y0 + hy - 1, x = x0 + hx - 1;
      const bool ok = y >= 0 && y < p.H && x >= 0 && x < p.W;
      const __half* src = xin;
      if (ok) {
        if (MODE == 1) {
          src = xin + (int64_t)y * p.x_y_stride + (int64_t)x * p.x_x_stride + c8 * 8;
        } else {          // conv channel (py, px, c): chunk c8 = (py * 2 + px) * (CS / 8) + c / 8 of source pixel (2y + py, 2x + px)
          constexpr int CS8 = Cfg::CH8 / 4;
          const int q = c8 / CS8, cc = c8 % CS8;
          src = xin + (int64_t)(2 * y + (q >> 1)) * p.x_y_stride + (int64_t)(2 * x + (q & 1)) * p.x_x_stride + cc * 8;
        }
      }
      const uint32_t dst = tile_s0 + hy * Cfg::ROWB + hx * Cfg::VS + ((c8 ^ Cfg::swz(hx)) << 4);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(ok ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;\n cp.async.wait_group 0;" ::: "memory");
  } else {
    constexpr int ITEMS = kC2HY * kC2HX * Cfg::CH8;
    constexpr int UNROLL = 4;
    for (int i0 = threadIdx.x; i0 < ITEMS; i0 += kC2Threads * UNROLL) {
      uint4 val[UNROLL];
      int dst[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const int idx = i0 + u * kC2Threads;
        const int pix = idx / Cfg::CH8, c8 = idx % Cfg::CH8;
        const int hy = pix / kC2HX, hx = pix - hy * kC2HX;
        const int y = y0 + hy - 1, x = x0 + hx - 1;
        val[u] = make_uint4(0u, 0u, 0u, 0u);
        dst[u] = idx < ITEMS ? hy * Cfg::ROWB + hx * Cfg::VS + ((c8 ^ Cfg::swz(hx)) << 4) : -1;
        if (idx < ITEMS && y >= 0 && y < p.H && x >= 0 && x < p.W) {
          // fp32 source through space-to-depth, converted while staging
          constexpr int CS8 = Cfg::CH8 / 4;                             // 16-byte fp16 chunks per SOURCE pixel
          const int q = c8 / CS8, cc = c8 % CS8;
          const float* src = reinterpret_cast<const float*>(p.x) + (int64_t)n * p.x_n_stride +
                             (int64_t)(2 * y + (q >> 1)) * p.x_y_stride + (int64_t)(2 * x + (q & 1)) * p.x_x_stride + cc * 8;
          const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src + 4));
          val[u].x = pack_half2_sat(a.x, a.y); val[u].y = pack_half2_sat(a.z, a.w);
          val[u].z = pack_half2_sat(b.x, b.y); val[u].w = pack_half2_sat(b.z, b.w);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (dst[u] >= 0) *reinterpret_cast<uint4*>(tile + dst[u]) = val[u];
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;
  const int mx = warp & 1, yb = (warp >> 1) * kC2WY;
  if (y0 + yb >= p.H || x0 + mx * 16 >= p.W) return;                     // warp-uniform; no barrier follows
  constexpr int CPL = 2 * NT;                                            // consecutive output channels per lane
  float acc[kC2WY][NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = t * CPL + nt * 2;
    const float b0 = p.bias ? __ldg(p.bias + c) : 0.f, b1 = p.bias ? __ldg(p.bias + c + 1) : 0.f;
#pragma unroll
    for (int oy = 0; oy < kC2WY; ++oy) { acc[oy][nt][0] = b0; acc[oy][nt][1] = b1; acc[oy][nt][2] = b0; acc[oy][nt][3] = b1; }
  }
#pragma unroll 1
  for (int j = 0; j < Cfg::KS; ++j) {
    const int dx = j / Cfg::KPD, part = j - dx * Cfg::KPD;
    uint2 bw[3][NT];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) bw[dy][nt] = __ldg(wfrag + ((dy * Cfg::KS + j) * NT + nt) * 32 + lane);
    const int v = mx * 16 + lrow + dx;
    const uint32_t aoff = tile_s + v * Cfg::VS + (((part * 2 + lhi) ^ Cfg::swz(v)) << 4) + yb * Cfg::ROWB;
#pragma unroll
    for (int py = 0; py < kC2WY + 2; ++py) {
      uint32_t a[4];
      ldmatrix_x4(a, aoff + py * Cfg::ROWB);
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int oy = py - dy;
        if (oy < 0 || oy >= kC2WY) continue;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) hmma16816(acc[oy][nt], a, bw[dy][nt].x, bw[dy][nt].y);
      }
    }
  }
  // ---- epilogue
  const int gx0 = x0 + mx * 16 + g, gx1 = gx0 + 8;
  if (FUSE) {
    // out = W1 * relu(conv) + b1 with the ReLU'd C fragments as A fragments: k-step kk takes n-tiles 2kk (k 0..7) and
    // 2kk + 1 (k 8..15); the host packs W1's rows in that (permuted) channel order and its columns permuted like above
    constexpr int NT1 = 4, KK1 = NT / 2, CPL1 = 2 * NT1;
    const uint2* w1 = reinterpret_cast<const uint2*>(p.wfrag1x1);
    uint2 b1w[KK1][NT1];
#pragma unroll
    for (int kk = 0; kk < KK1; ++kk)
#pragma unroll
      for (int nt = 0; nt < NT1; ++nt) b1w[kk][nt] = __ldg(w1 + (kk * NT1 + nt) * 32 + lane);
    float bias1[NT1][2];
#pragma unroll
    for (int nt = 0; nt < NT1; ++nt) {
      bias1[nt][0] = p.bias1x1 ? __ldg(p.bias1x1 + t * CPL1 + nt * 2) : 0.f;
      bias1[nt][1] = p.bias1x1 ? __ldg(p.bias1x1 + t * CPL1 + nt * 2 + 1) : 0.f;
    }
    float* out = reinterpret_cast<float*>(p.out) + (int64_t)n * p.o_n_stride;
#pragma unroll
    for (int oy = 0; oy < kC2WY; ++oy) {
      const int gy = y0 + yb + oy;
      float o[NT1][4];
#pragma unroll
      for (int nt = 0; nt < NT1; ++nt) { o[nt][0] = bias1[nt][0]; o[nt][1] = bias1[nt][1]; o[nt][2] = bias1[nt][0]; o[nt][3] = bias1[nt][1]; }
#pragma unroll
      for (int kk = 0; kk < KK1; ++kk) {
        uint32_t a[4];
        const float* c0 = acc[oy][2 * kk];
        const float* c1 = acc[oy][2 * kk + 1];
        const float lo = p.relu ? 0.f : -3.4e38f;
        a[0] = pack_half2_sat(fmaxf(c0[0], lo), fmaxf(c0[1], lo)); a[1] = pack_half2_sat(fmaxf(c0[2], lo), fmaxf(c0[3], lo));
        a[2] = pack_half2_sat(fmaxf(c1[0], lo), fmaxf(c1[1], lo)); a[3] = pack_half2_sat(fmaxf(c1[2], lo), fmaxf(c1[3], lo));
#pragma unroll
        for (int nt = 0; nt < NT1; ++nt) hmma16816(o[nt], a, b1w[kk][nt].x, b1w[kk][nt].y);
      }
      if (gy >= p.H) continue;
      float* orow = out + (int64_t)gy * p.o_y_stride + t * CPL1;
      if (gx0 < p.W) {
        float* q = orow + (int64_t)gx0 * p.o_x_stride;
        *reinterpret_cast<float4*>(q) = make_float4(o[0][0], o[0][1], o[1][0], o[1][1]);
        *reinterpret_cast<float4*>(q + 4) = make_float4(o[2][0], o[2][1], o[3][0], o[3][1]);
      }
      if (gx1 < p.W) {
        float* q = orow + (int64_t)gx1 * p.o_x_stride;
        *reinterpret_cast<float4*>(q) = make_float4(o[0][2], o[0][3], o[1][2], o[1][3]);
        *reinterpret_cast<float4*>(q + 4) = make_float4(o[2][2], o[2][3], o[3][2], o[3][3]);
      }
    }
    return;
  }
#pragma unroll
  for (int oy = 0; oy < kC2WY; ++oy) {
    const int gy = y0 + yb + oy;
    if (gy >= p.H) continue;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int gx = e ? gx1 : gx0;
      if (gx >= p.W) continue;
      float v[CPL];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        v[2 * nt] = acc[oy][nt][2 * e]; v[2 * nt + 1] = acc[oy][nt][2 * e + 1];
        if (p.relu) { v[2 * nt] = fmaxf(v[2 * nt], 0.f); v[2 * nt + 1] = fmaxf(v[2 * nt + 1], 0.f); }
      }
      if (p.out_half) {
        __half* q = reinterpret_cast<__half*>(p.out) + (int64_t)n * p.o_n_stride + (int64_t)gy * p.o_y_stride + (int64_t)gx * p.o_x_stride + t * CPL;
        if constexpr (CPL == 4) {
          *reinterpret_cast<uint2*>(q) = make_uint2(pack_half2_sat(v[0], v[1]), pack_half2_sat(v[2], v[3]));
        } else {
          *reinterpret_cast<uint4*>(q) = make_uint4(pack_half2_sat(v[0], v[1]), pack_half2_sat(v[2], v[3]), pack_half2_sat(v[CPL - 4], v[CPL - 3]),
                                                     pack_half2_sat(v[CPL - 2], v[CPL - 1]));
        }
      } else {
        float* q = reinterpret_cast<float*>(p.out) + (int64_t)n * p.o_n_stride + (int64_t)gy * p.o_y_stride + (int64_t)gx * p.o_x_stride + t * CPL;
#pragma unroll
        for (int i = 0; i < CPL; i += 4) *reinterpret_cast<float4*>(q + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
    }
  }
}

template <int CIN, int NT, int MODE, bool FUSE, int TY>
static int launch_c2t(const bmv_conv2d_params& p, cudaStream_t st) {
  using Cfg = C2Cfg<CIN, TY>;
  constexpr int kC2Threads = Cfg::THREADS, kC2TY = TY;
  const size_t smem = (size_t)Cfg::TILE_BYTES;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(conv2d_k3_mma_kernel<CIN, NT, MODE, FUSE, TY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv2d_k3_mma_kernel<CIN, NT, MODE, FUSE, TY>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("bmv_conv2d_k3: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const dim3 grid((unsigned)(((p.W + kC2TX - 1) / kC2TX) * ((p.H + kC2TY - 1) / kC2TY)), (unsigned)p.N);
  conv2d_k3_mma_kernel<CIN, NT, MODE, FUSE, TY><<<grid, kC2Threads, smem, st>>>(p);
  return check_launch("bmv_conv2d_k3");
}

// 16-row tiles while they fill the GPU at least ~3 times over, 8-row tiles (twice as many CTAs of half the size) below that
template <int CIN, int NT, int MODE, bool FUSE>
static int launch_c2(const bmv_conv2d_params& p, cudaStream_t st) {
  static const int ty_env = getenv("BMV_C2_TY") ? atoi(getenv("BMV_C2_TY")) : 0;   // measurements
  const int64_t tiles16 = (int64_t)((p.W + kC2TX - 1) / kC2TX) * ((p.H + 15) / 16) * p.N;
  const bool small = ty_env ? ty_env == 8 : tiles16 < 900;
  return small ? launch_c2t<CIN, NT, MODE, FUSE, 8>(p, st) : launch_c2t<CIN, NT, MODE, FUSE, 16>(p, st);
}

}  // namespace bmv

extern "C" BMV_API int bmv_conv2d_k3(const bmv_conv2d_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_conv2d_k3");
  using namespace bmv;
  BMV_REQUIRE(p && p->x && p->wfrag && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_conv2d_k3: null pointer");
  BMV_REQUIRE(p->N >= 1 && p->N <= 65535 && p->H >= 1 && p->W >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_conv2d_k3: bad size");
  const int m = p->in_half ? 8 : 4;
  BMV_REQUIRE(p->x_n_stride % m == 0 && p->x_y_stride % m == 0 && p->x_x_stride % m == 0 && ((uintptr_t)p->x & 15) == 0 &&
                  ((uintptr_t)p->wfrag & 15) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_conv2d_k3: input must be channels-last with 16-byte aligned pixels");
  const bool fuse = p->wfrag1x1 != nullptr;
  const int co = fuse ? p->C1x1_out : p->Cout;                           // channels of the tensor that is written
  const int om = (p->out_half && !fuse) ? 8 : 4;
  BMV_REQUIRE(p->o_n_stride % om == 0 && p->o_y_stride % om == 0 && p->o_x_stride % om == 0 && ((uintptr_t)p->out & 15) == 0 && co % 8 == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_conv2d_k3: output must be channels-last with 16-byte aligned pixels");
  BMV_REQUIRE(!fuse || (!p->out_half && p->C1x1_out == 32 && p->Cout == 32 && ((uintptr_t)p->wfrag1x1 & 7) == 0), BMV_ERR_UNSUPPORTED_SHAPE,
              "bmv_conv2d_k3: the fused 1x1 layer is instantiated for 32 -> 32 channels, fp32 output");
  cudaStream_t st = (cudaStream_t)stream;
  const int mode = p->s2d ? (p->in_half ? 3 : 2) : 1;
  BMV_REQUIRE(mode != 1 || p->in_half, BMV_ERR_UNSUPPORTED_SHAPE, "bmv_conv2d_k3: the dense mode reads fp16");
  if (mode == 2 && p->Cin == 32 && p->Cout == 16 && !fuse) return launch_c2<32, 2, 2, false>(*p, st);
  if (mode == 2 && p->Cin == 64 && p->Cout == 32 && !fuse) return launch_c2<64, 4, 2, false>(*p, st);
  if (mode == 3 && p->Cin == 32 && p->Cout == 16 && !fuse) return launch_c2<32, 2, 3, false>(*p, st);
  if (mode == 3 && p->Cin == 64 && p->Cout == 32 && !fuse) return launch_c2<64, 4, 3, false>(*p, st);
  if (mode == 1 && p->Cin == 16 && p->Cout == 16 && !fuse) return launch_c2<16, 2, 1, false>(*p, st);
  if (mode == 1 && p->Cin == 32 && p->Cout == 32 && !fuse) return launch_c2<32, 4, 1, false>(*p, st);
  if (mode == 1 && p->Cin == 32 && p->Cout == 32 && fuse) return launch_c2<32, 4, 1, true>(*p, st);
  set_error("bmv_conv2d_k3: (Cin=%d, Cout=%d, s2d=%d, fused 1x1=%d) not instantiated", p->Cin, p->Cout, p->s2d, (int)fuse);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}

// words (uint32) of the fragment-ordered 3x3 weight buffer, -1 if (Cin, Cout) is not instantiated
extern "C" BMV_API int bmv_conv2d_k3_weight_words(int Cin, int Cout) {
  if ((Cin == 32 && Cout == 16) || (Cin == 64 && Cout == 32) || (Cin == 16 && Cout == 16) || (Cin == 32 && Cout == 32))
    return 3 * (3 * Cin / 16) * (Cout / 8) * 32 * 2;
  return -1;
}
