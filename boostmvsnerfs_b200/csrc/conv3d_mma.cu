// Tensor-core 3x3x3 / stride-1 / pad-1 convolution for the FULL-RESOLUTION layers of the kept 3-D
// cost regularisers (conv0: C->8, and the merged output heads 8->9), channels-last fp32 in and out.
//
// Why (north_star: "unless ncu shows a hand-written tensor-core tile pays off"): with every other stage
// optimised these cuDNN layers dominate the frame and run at 3 % of the TF32 peak AND 6 % of the HBM
// bandwidth (profiles/round1_fpn_layers.md: cost_reg_1.conv0 0.93 ms for 14.4 GMAC / 401 MB; heads
// 0.6 ms) — cuDNN has no good kernel for 8..32-channel 3-D convolutions.
//
// Precision contract: operands are rounded to fp16 (11-bit significand), accumulation is fp32 — the
// same class as cuDNN's TF32 path (10+1 bits), which is what PyTorch uses for convolutions by default
// (torch.backends.cudnn.allow_tf32 = True).  The host side (inference_plan.py) therefore routes a layer
// here ONLY when allow_tf32 is True; with TF32 disabled (the strict 1e-4 parity tests) cuDNN's fp32
// path runs instead.
//
// Implicit GEMM on mma.sync.m16n8k16: M = 16 consecutive x voxels, N = 8 output channels per n-tile,
// K = one 32-byte segment of a staged input row (a "k-step"):
//     Cin 16: the 16 channels of voxel x+dx                         (3 k-steps per (dz,dy))
//     Cin 32: 16 of the 32 channels of voxel x+dx                   (6 k-steps)
//     Cin  8: the 8+8 channels of voxels x+dx, x+dx+1 — two taps per MMA; the third tap is paired with
//             zero weights                                          (2 k-steps)
// A CTA stages an (8+2)x(TH+2)x(32+2) input tile as fp16 in shared memory (zero outside the volume = the
// convolution's padding).  A warp owns WD x TH x 16 output voxels and walks the INPUT rows: each A
// fragment (one ldmatrix.x4) is used for every (dz,dy) whose output row the warp owns, so shared-memory
// traffic per MMA drops ~3x against a per-output-tile loop; weights are host-arranged B fragments read from
// shared memory (register-resident weights cost a CTA of occupancy: 151 vs 145 us on conv0 of level 1, so the
// BREG switch of ConvCfg is off everywhere).  The input may be fp32 (converted while staging, one warp per
// staged row) or fp16 (K1 emits the cost volume in fp16 for this kernel: the staged tile is a straight copy).
// Bias + ReLU are fused in the epilogue; channels >= `split` can go to a second tensor (the depth logits of
// the merged heads).
#include <cuda.h>

#include <cstring>
#include <cuda_fp16.h>

#include "bmv_internal.cuh"
#include "conv_mma.cuh"

namespace bmv {

constexpr int kConvThreads = 256;
constexpr int kConvWarps = kConvThreads / 32;

template <int CIN, int NTILES> struct ConvCfg;
// Shared-memory voxel v (index in its staged row) holds its 16-byte chunk c at chunk c ^ swz(v): with
// 32- and 64-byte voxels this makes the 8 rows of every ldmatrix 8x8 block hit 8 distinct 16-byte bank groups.
template <int NTILES> struct ConvCfg<16, NTILES> {
  static constexpr int VS = 32, KS = 3, NT = NTILES, TH = 4, WD = NTILES == 1 ? 2 : 1, EXTRA = 0;
  static constexpr bool BREG = false;
  __device__ static __forceinline__ int swz(int v) { return (v >> 2) & 1; }
  // k-step j: voxel offset and first 16-byte chunk of the lane's row segment (hi = lane / 16)
  __device__ static __forceinline__ int step_voxel(int j, int hi) { return j; }
  __device__ static __forceinline__ int step_chunk(int j, int hi) { return hi; }
};
template <> struct ConvCfg<32, 1> {
  static constexpr int VS = 64, KS = 6, NT = 1, TH = 2, WD = 2, EXTRA = 0;
  static constexpr bool BREG = false;
  __device__ static __forceinline__ int swz(int v) { return (v >> 1) & 3; }
  __device__ static __forceinline__ int step_voxel(int j, int hi) { return j >> 1; }
  __device__ static __forceinline__ int step_chunk(int j, int hi) { return (j & 1) * 2 + hi; }
};
template <> struct ConvCfg<8, 2> {
  static constexpr int VS = 16, KS = 2, NT = 2, TH = 4, WD = 1, EXTRA = 1;   // +1 voxel: the unpaired tap reads x+3
  static constexpr bool BREG = false;
  __device__ static __forceinline__ int swz(int v) { return 0; }
  __device__ static __forceinline__ int step_voxel(int j, int hi) { return 2 * j + hi; }
  __device__ static __forceinline__ int step_chunk(int j, int hi) { return 0; }
};

template <int CIN, int NTILES>
struct ConvTile {
  using Cfg = ConvCfg<CIN, NTILES>;
  static constexpr int TD = 8, TW = 32, TH = Cfg::TH;
  static constexpr int HD = TD + 2, HH = TH + 2, HW = TW + 2;
  static constexpr int ROWV = HW + Cfg::EXTRA;                          // staged voxels per row
  static constexpr int ROWB = ROWV * Cfg::VS;                           // bytes per staged row
  static constexpr int TILE_BYTES = HD * HH * ROWB;
  static constexpr int W_WORDS = 9 * Cfg::KS * Cfg::NT * 32 * 2;
  static constexpr int JOBS = (TD / Cfg::WD) * (TW / 16);
};

// ---- TMA staging (fp16 input, Cin 16 / 32): ONE elected thread issues a 5-D tiled bulk-tensor copy of the whole
// (10 x (TH+2) x 34 x Cin) halo tile; out-of-bounds coordinates (negative or past the volume) are zero-filled by the
// TMA unit — that IS the convolution's padding — and the hardware 32B / 64B swizzle gives the same conflict-free
// ldmatrix pattern as the manual swizzle of the register-staged path.  All threads wait on an mbarrier.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spins = 0; !done; ++spins) {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    if (spins > (1u << 24)) __trap();                                   // a faulted copy must not hang the GPU
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

template <int CIN, int NTILES, bool IN_HALF, bool TMA>
__global__ void __launch_bounds__(kConvThreads, (CIN == 8 || (CIN == 16 && NTILES == 1)) ? 3 : 2)
conv3d_k3_mma_kernel(bmv_conv3d_params p, const __grid_constant__ CUtensorMap tmap) {
  using T = ConvTile<CIN, NTILES>;
  using Cfg = ConvCfg<CIN, NTILES>;
  constexpr int NT = Cfg::NT, KS = Cfg::KS, WD = Cfg::WD, TH = T::TH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // the TMA destination must be 128-byte aligned: round the dynamic window up (the launcher allocates 128 B extra)
  unsigned char* smem = smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u);
  unsigned char* tile = smem;
  const uint2* wfrag = reinterpret_cast<const uint2*>(smem + T::TILE_BYTES);
  __shared__ __align__(8) uint64_t s_mbar;
  // ---- which output tile
  const int tiles_w = (p.W + T::TW - 1) / T::TW, tiles_h = (p.H + TH - 1) / TH, tiles_d = (p.D + T::TD - 1) / T::TD;
  int b = blockIdx.x;
  const int tw = b % tiles_w; b /= tiles_w;
  const int th = b % tiles_h; b /= tiles_h;
  const int td = b % tiles_d; b /= tiles_d;
  const int n = b;
  const int x0 = tw * T::TW, y0 = th * TH, d0 = td * T::TD;
  if (TMA) {
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(mbar, (uint32_t)T::TILE_BYTES);
      tma_load_5d((uint32_t)__cvta_generic_to_shared(tile), &tmap, mbar, 0, x0 - 1, y0 - 1, d0 - 1, n);
    }
  }
  // ---- weights -> shared memory (overlaps the bulk copy on the TMA path)
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wfrag);
    uint4* dst = reinterpret_cast<uint4*>(smem + T::TILE_BYTES);
    for (int i = threadIdx.x; i < T::W_WORDS / 4; i += kConvThreads) dst[i] = __ldg(src + i);
  }
  if (TMA) {
    mbar_wait((uint32_t)__cvta_generic_to_shared(&s_mbar), 0);
  } else if (IN_HALF) {
    // ---- fp16 input: the staged tile is a straight copy, 16 bytes (8 channels) per lane per pass
    const __half* xin = reinterpret_cast<const __half*>(p.x) + (int64_t)n * p.x_n_stride;
    constexpr int CH8 = CIN / 8;                                        // 16-byte chunks per voxel
    constexpr int PER_ROW = T::ROWV * CH8;
    constexpr int ROWS = T::HD * T::HH;
    constexpr int P = (PER_ROW + 31) / 32;
    constexpr int VPP = 32 / CH8;
    constexpr int RB = 4;                                               // rows in flight per warp
    const int sl = threadIdx.x & 31, sw = threadIdx.x >> 5;
    const int c8 = sl % CH8, hx0 = sl / CH8;
    const int64_t lane_off = (int64_t)(x0 - 1 + hx0) * p.x_x_stride + c8 * 8, pass_off = (int64_t)VPP * p.x_x_stride;
    for (int row = sw; row < ROWS; row += RB * kConvWarps) {
      uint4 val[RB][P];
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int r = row + rr * kConvWarps;
        const int hd = r / T::HH, hy = r - hd * T::HH;
        const int gy = y0 + hy - 1, gd = d0 + hd - 1;
        const bool row_ok = r < ROWS && gy >= 0 && gy < p.H && gd >= 0 && gd < p.D;
        const __half* src = xin + (int64_t)gd * p.x_d_stride + (int64_t)gy * p.x_y_stride + lane_off;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int hx = hx0 + VPP * k, gx = x0 - 1 + hx;
          val[rr][k] = make_uint4(0u, 0u, 0u, 0u);
          if (row_ok && hx < T::HW && gx >= 0 && gx < p.W) val[rr][k] = __ldg(reinterpret_cast<const uint4*>(src + k * pass_off));
        }
      }
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int r = row + rr * kConvWarps;
        if (r < ROWS) {
#pragma unroll
          for (int k = 0; k < P; ++k) {
            const int hx = hx0 + VPP * k;
            if (hx < T::ROWV) *reinterpret_cast<uint4*>(tile + r * T::ROWB + hx * Cfg::VS + ((c8 ^ Cfg::swz(hx)) << 4)) = val[rr][k];
          }
        }
      }
    }
  } else
  // ---- stage the input tile with halo as fp16.  One warp per staged row (a contiguous run of float4 in
  // channels-last memory): lane l covers float4 l, l+32, ... of the row, so its channel chunk c4 is fixed,
  // its voxel advances by 32/CH4 per pass, and everything row-dependent is warp-uniform.
  {
    const float* xin = p.x + (int64_t)n * p.x_n_stride;
    constexpr int CH4 = CIN / 4;                                        // float4 chunks per voxel (power of two)
    constexpr int PER_ROW = T::ROWV * CH4;
    constexpr int ROWS = T::HD * T::HH;
    constexpr int P = (PER_ROW + 31) / 32;                              // passes per row
    constexpr int VPP = 32 / CH4;                                       // voxels per pass
    constexpr int RB = P >= 8 ? 1 : 2;                                  // rows in flight per warp
    const int sl = threadIdx.x & 31, sw = threadIdx.x >> 5;
    const int c4 = sl % CH4, hx0 = sl / CH4;
    const int64_t lane_off = (int64_t)(x0 - 1 + hx0) * p.x_x_stride + c4 * 4, pass_off = (int64_t)VPP * p.x_x_stride;
    for (int row = sw; row < ROWS; row += RB * kConvWarps) {
      float4 val[RB][P];
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int r = row + rr * kConvWarps;
        const int hd = r / T::HH, hy = r - hd * T::HH;
        const int gy = y0 + hy - 1, gd = d0 + hd - 1;
        const bool row_ok = r < ROWS && gy >= 0 && gy < p.H && gd >= 0 && gd < p.D;
        const float* src = xin + (int64_t)gd * p.x_d_stride + (int64_t)gy * p.x_y_stride + lane_off;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int hx = hx0 + VPP * k, gx = x0 - 1 + hx;
          val[rr][k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && hx < T::HW && gx >= 0 && gx < p.W) val[rr][k] = __ldg(reinterpret_cast<const float4*>(src + k * pass_off));
        }
      }
#pragma unroll
      for (int rr = 0; rr < RB; ++rr) {
        const int r = row + rr * kConvWarps;
        if (r < ROWS) {
#pragma unroll
          for (int k = 0; k < P; ++k) {
            const int hx = hx0 + VPP * k;
            if (hx < T::ROWV) {
              uint2 pk;
              pk.x = pack_half2_sat(val[rr][k].x, val[rr][k].y);
              pk.y = pack_half2_sat(val[rr][k].z, val[rr][k].w);
              *reinterpret_cast<uint2*>(tile + r * T::ROWB + hx * Cfg::VS + (((c4 >> 1) ^ Cfg::swz(hx)) << 4) + (c4 & 1) * 8) = pk;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  // ldmatrix row of this lane: matrix m = lane/8 -> rows (m&1)*8 + lane%8; m>>1 selects the upper 8 k (next 16 bytes)
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;
  // ---- weights to registers when they fit
  uint2 breg[Cfg::BREG ? 9 * KS * NT : 1];
  if (Cfg::BREG) {
#pragma unroll
    for (int i = 0; i < 9 * KS * NT; ++i) breg[i] = wfrag[i * 32 + lane];
  }
  // the input may be stored pre-multiplied by a power of two (bmv_volume_scale): start from s*bias, undo with 1/s
  const float in_sc = p.in_scale ? __ldg(p.in_scale) : 1.f, in_isc = p.in_scale ? __ldg(p.in_scale + 1) : 1.f;
  float bias0[NT], bias1[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int c = nt * 8 + 2 * t;
    bias0[nt] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) * in_sc : 0.f;
    bias1[nt] = (p.bias && c + 1 < p.Cout) ? __ldg(p.bias + c + 1) * in_sc : 0.f;
  }
  float* out = p.out + (int64_t)n * p.o_n_stride;
  float* out2 = p.out2 ? p.out2 + (int64_t)n * p.o2_n_stride : nullptr;
  const int split = p.out2 ? p.split : p.Cout;
  const bool vec_ok = (p.o_x_stride % 2 == 0) && (p.o_y_stride % 2 == 0) && (p.o_d_stride % 2 == 0) &&
                      (p.o_n_stride % 2 == 0) && (((uintptr_t)p.out & 7) == 0);

  // ---- each warp owns WD x TH output rows of 16 voxels and walks the input rows they touch
  for (int job = warp; job < T::JOBS; job += kConvWarps) {
    const int mx = job % (T::TW / 16), db = (job / (T::TW / 16)) * WD;
    if (d0 + db >= p.D || x0 + mx * 16 >= p.W) continue;                // warp-uniform
    float acc[WD][TH][NT][4];
#pragma unroll
    for (int od = 0; od < WD; ++od)
#pragma unroll
      for (int oy = 0; oy < TH; ++oy)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          acc[od][oy][nt][0] = bias0[nt]; acc[od][oy][nt][1] = bias1[nt];
          acc[od][oy][nt][2] = bias0[nt]; acc[od][oy][nt][3] = bias1[nt];
        }
    uint32_t aoff[KS];                                                  // this lane's byte offset in a staged row, per k-step
#pragma unroll
    for (int j = 0; j < KS; ++j) {
      const int v = mx * 16 + lrow + Cfg::step_voxel(j, lhi);
      aoff[j] = tile_s + v * Cfg::VS + ((Cfg::step_chunk(j, lhi) ^ Cfg::swz(v)) << 4);
    }
#pragma unroll
    for (int pd = 0; pd < WD + 2; ++pd) {
#pragma unroll
      for (int py = 0; py < TH + 2; ++py) {
#pragma unroll
        for (int j = 0; j < KS; ++j) {
          uint32_t a[4];
          if (TMA) {
            // dense [d][y][x][c] box written by the TMA unit with its 32B / 64B swizzle: 16-byte chunk index XOR
            // address bits [7] (32B voxels) / [8:7] (64B voxels); the tile base is 1024-byte aligned
            const uint32_t lin = tile_s + ((((db + pd) * T::HH + py) * T::ROWV) + mx * 16 + lrow + Cfg::step_voxel(j, lhi)) * Cfg::VS +
                                 (Cfg::step_chunk(j, lhi) << 4);
            ldmatrix_x4(a, lin ^ (((lin >> 7) & (Cfg::VS == 32 ? 1u : (Cfg::VS == 64 ? 3u : 0u))) << 4));
          } else {
            ldmatrix_x4(a, aoff[j] + ((db + pd) * T::HH + py) * T::ROWB);
          }
#pragma unroll
          for (int dz = 0; dz < 3; ++dz) {
            const int od = pd - dz;
            if (od < 0 || od >= WD) continue;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const int oy = py - dy;
              if (oy < 0 || oy >= TH) continue;
#pragma unroll
              for (int nt = 0; nt < NT; ++nt) {
                const int wi = ((dz * 3 + dy) * KS + j) * NT + nt;
                const uint2 bw = Cfg::BREG ? breg[Cfg::BREG ? wi : 0] : wfrag[wi * 32 + lane];
                hmma16816(acc[od][oy][nt], a, bw.x, bw.y);
              }
            }
          }
        }
      }
    }
    // ---- epilogue: ReLU, store fp32 channels-last
    const int gx0 = x0 + mx * 16 + g, gx1 = gx0 + 8;
#pragma unroll
    for (int od = 0; od < WD; ++od) {
      const int gd = d0 + db + od;
      if (gd >= p.D) continue;
#pragma unroll
      for (int oy = 0; oy < TH; ++oy) {
        const int gy = y0 + oy;
        if (gy >= p.H) continue;
        float* orow = out + (int64_t)gd * p.o_d_stride + (int64_t)gy * p.o_y_stride;
        float* orow2 = out2 ? out2 + (int64_t)gd * p.o2_d_stride + (int64_t)gy * p.o2_y_stride : nullptr;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          const int c = nt * 8 + 2 * t;
          float v0 = acc[od][oy][nt][0] * in_isc, v1 = acc[od][oy][nt][1] * in_isc, v2 = acc[od][oy][nt][2] * in_isc, v3 = acc[od][oy][nt][3] * in_isc;
          if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
          if (p.out_half) {                                             // fp16 storage (launcher: even Cout, no split)
            __half* hrow = reinterpret_cast<__half*>(p.out) + (int64_t)n * p.o_n_stride + (int64_t)gd * p.o_d_stride +
                           (int64_t)gy * p.o_y_stride;
            if (c < p.Cout) {
              if (gx0 < p.W) *reinterpret_cast<uint32_t*>(hrow + (int64_t)gx0 * p.o_x_stride + c) = pack_half2_sat(v0, v1);
              if (gx1 < p.W) *reinterpret_cast<uint32_t*>(hrow + (int64_t)gx1 * p.o_x_stride + c) = pack_half2_sat(v2, v3);
            }
          } else if (vec_ok && nt * 8 + 8 <= split) {                   // whole n-tile lands in `out`: 8-byte stores
            if (gx0 < p.W) *reinterpret_cast<float2*>(orow + (int64_t)gx0 * p.o_x_stride + c) = make_float2(v0, v1);
            if (gx1 < p.W) *reinterpret_cast<float2*>(orow + (int64_t)gx1 * p.o_x_stride + c) = make_float2(v2, v3);
          } else {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ce = c + e;
              if (ce >= p.Cout) continue;
              float* r0 = ce < split ? orow : orow2;
              const int64_t xs = ce < split ? p.o_x_stride : p.o2_x_stride;
              const int cc = ce < split ? ce : ce - split;
              if (gx0 < p.W) r0[(int64_t)gx0 * xs + cc] = e ? v1 : v0;
              if (gx1 < p.W) r0[(int64_t)gx1 * xs + cc] = e ? v3 : v2;
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------ stride 2 (8 -> 16 channels: conv1 of the U-Nets)
// out[o] = sum_k x[2o - 1 + k] w[k].  Same implicit GEMM; the A rows of an M-tile are every second staged
// voxel (ldmatrix takes one address per row), taps kx = 0,1 are the contiguous voxel pair (2o, 2o+1) of the
// staged row, kx = 2 is paired with zero weights.  CTA: 4 x 4 x 32 outputs from a 9 x 9 x 66 input tile.
struct ConvS2 {
  static constexpr int TD = 4, TH = 2, TW = 32, NT = 2, KS = 2;
  static constexpr int HD = 2 * TD + 1, HH = 2 * TH + 1, ROWV = 2 * TW + 2;
  static constexpr int ROWB = ROWV * 16;
  static constexpr int TILE_BYTES = HD * HH * ROWB;
  static constexpr int W_WORDS = 9 * KS * NT * 32 * 2;
};

template <bool TMA>
__global__ void __launch_bounds__(kConvThreads, 4)
conv3d_k3s2_c8_mma_kernel(bmv_conv3d_params p, int Do, int Ho, int Wo, const __grid_constant__ CUtensorMap tmap) {
  using T = ConvS2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((128u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 127u)) & 127u);
  unsigned char* tile = smem;
  const uint2* wfrag = reinterpret_cast<const uint2*>(smem + T::TILE_BYTES);
  __shared__ __align__(8) uint64_t s_mbar;
  const int tiles_w = (Wo + T::TW - 1) / T::TW, tiles_h = (Ho + T::TH - 1) / T::TH, tiles_d = (Do + T::TD - 1) / T::TD;
  int b = blockIdx.x;
  const int tw = b % tiles_w; b /= tiles_w;
  const int th = b % tiles_h; b /= tiles_h;
  const int td = b % tiles_d; b /= tiles_d;
  const int n = b;
  const int x0 = tw * T::TW, y0 = th * T::TH, d0 = td * T::TD;          // output coordinates
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (TMA) {                                                            // fp16 input: one bulk-tensor copy of the 9 x 9 x 66 tile
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&s_mbar);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(mbar, (uint32_t)T::TILE_BYTES);
      tma_load_5d((uint32_t)__cvta_generic_to_shared(tile), &tmap, mbar, 0, 2 * x0 - 1, 2 * y0 - 1, 2 * d0 - 1, n);
    }
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wfrag);
    uint4* dst = reinterpret_cast<uint4*>(smem + T::TILE_BYTES);
    for (int i = threadIdx.x; i < T::W_WORDS / 4; i += kConvThreads) dst[i] = __ldg(src + i);
  }
  if (TMA) {
    mbar_wait((uint32_t)__cvta_generic_to_shared(&s_mbar), 0);
  } else {
    const float* xin = p.x + (int64_t)n * p.x_n_stride;
    constexpr int PER_ROW = T::ROWV * 2, ROWS = T::HD * T::HH, P = (PER_ROW + 31) / 32, VPP = 16;
    const int c4 = lane & 1, hx0 = lane >> 1;
    const int64_t lane_off = (int64_t)(2 * x0 - 1 + hx0) * p.x_x_stride + c4 * 4, pass_off = (int64_t)VPP * p.x_x_stride;
    for (int row = warp; row < ROWS; row += 2 * kConvWarps) {
      float4 val[2][P];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = row + rr * kConvWarps;
        const int hd = r / T::HH, hy = r - hd * T::HH;
        const int gy = 2 * y0 - 1 + hy, gd = 2 * d0 - 1 + hd;
        const bool row_ok = r < ROWS && gy >= 0 && gy < p.H && gd >= 0 && gd < p.D;
        const float* src = xin + (int64_t)gd * p.x_d_stride + (int64_t)gy * p.x_y_stride + lane_off;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int hx = hx0 + VPP * k, gx = 2 * x0 - 1 + hx;
          val[rr][k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok && hx < T::ROWV && gx >= 0 && gx < p.W) val[rr][k] = __ldg(reinterpret_cast<const float4*>(src + k * pass_off));
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int r = row + rr * kConvWarps;
        if (r < ROWS) {
#pragma unroll
          for (int k = 0; k < P; ++k) {
            const int hx = hx0 + VPP * k;
            if (hx < T::ROWV) *reinterpret_cast<uint2*>(tile + r * T::ROWB + hx * 16 + c4 * 8) = pack_half4(val[rr][k]);
          }
        }
      }
    }
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;
  float bias0[T::NT], bias1[T::NT];
#pragma unroll
  for (int nt = 0; nt < T::NT; ++nt) {
    const int c = nt * 8 + 2 * t;
    bias0[nt] = (p.bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
    bias1[nt] = (p.bias && c + 1 < p.Cout) ? __ldg(p.bias + c + 1) : 0.f;
  }
  float* out = p.out + (int64_t)n * p.o_n_stride;
  // one job per warp: output plane od, x half mx, all TH rows
  {
    const int mx = warp & 1, od = warp >> 1;
    if (d0 + od < Do && x0 + mx * 16 < Wo) {
      float acc[T::TH][T::NT][4];
#pragma unroll
      for (int oy = 0; oy < T::TH; ++oy)
#pragma unroll
        for (int nt = 0; nt < T::NT; ++nt) {
          acc[oy][nt][0] = bias0[nt]; acc[oy][nt][1] = bias1[nt]; acc[oy][nt][2] = bias0[nt]; acc[oy][nt][3] = bias1[nt];
        }
      const uint32_t lane_base = tile_s + (2 * (mx * 16 + lrow) + lhi) * 16;
#pragma unroll
      for (int dz = 0; dz < 3; ++dz) {
#pragma unroll
        for (int py = 0; py < 2 * T::TH + 1; ++py) {
#pragma unroll
          for (int j = 0; j < T::KS; ++j) {
            uint32_t a[4];
            ldmatrix_x4(a, lane_base + ((2 * od + dz) * T::HH + py) * T::ROWB + j * 32);
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              if ((py - dy) < 0 || ((py - dy) & 1) || (py - dy) / 2 >= T::TH) continue;
              const int oy = (py - dy) / 2;
#pragma unroll
              for (int nt = 0; nt < T::NT; ++nt) {
                const uint2 bw = wfrag[((((dz * 3 + dy) * T::KS + j) * T::NT + nt) << 5) + lane];
                hmma16816(acc[oy][nt], a, bw.x, bw.y);
              }
            }
          }
        }
      }
      const int gx0 = x0 + mx * 16 + g, gx1 = gx0 + 8;
#pragma unroll
      for (int oy = 0; oy < T::TH; ++oy) {
        const int gy = y0 + oy;
        if (gy >= Ho) continue;
        float* orow = out + (int64_t)(d0 + od) * p.o_d_stride + (int64_t)gy * p.o_y_stride;
#pragma unroll
        for (int nt = 0; nt < T::NT; ++nt) {
          const int c = nt * 8 + 2 * t;
          if (c + 1 >= p.Cout + 1 || c >= p.Cout) continue;
          float v0 = acc[oy][nt][0], v1 = acc[oy][nt][1], v2 = acc[oy][nt][2], v3 = acc[oy][nt][3];
          if (p.relu) { v0 = fmaxf(v0, 0.f); v1 = fmaxf(v1, 0.f); v2 = fmaxf(v2, 0.f); v3 = fmaxf(v3, 0.f); }
          if (p.out_half) {
            __half* hrow = reinterpret_cast<__half*>(p.out) + (int64_t)n * p.o_n_stride + (int64_t)(d0 + od) * p.o_d_stride +
                           (int64_t)gy * p.o_y_stride;
            if (gx0 < Wo) *reinterpret_cast<uint32_t*>(hrow + (int64_t)gx0 * p.o_x_stride + c) = pack_half2_sat(v0, v1);
            if (gx1 < Wo) *reinterpret_cast<uint32_t*>(hrow + (int64_t)gx1 * p.o_x_stride + c) = pack_half2_sat(v2, v3);
          } else {
            if (gx0 < Wo) *reinterpret_cast<float2*>(orow + (int64_t)gx0 * p.o_x_stride + c) = make_float2(v0, v1);
            if (gx1 < Wo) *reinterpret_cast<float2*>(orow + (int64_t)gx1 * p.o_x_stride + c) = make_float2(v2, v3);
          }
        }
      }
    }
  }
}

static thread_local int g_last_conv3d_tma = 0;      // which staging path the last stride-1 launch of this thread used (tests)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

// 5-D map (C, W, H, D, N) of the fp16 channels-last input, box = one halo tile; false if the layout does not qualify
template <int CIN, int NTILES>
static bool make_input_map(const bmv_conv3d_params& p, CUtensorMap* map) {
  using T = ConvTile<CIN, NTILES>;
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || !p.in_half) return false;
  if (p.x_x_stride != CIN) return false;                                // voxels contiguous along x
  const cuuint64_t dims[5] = {(cuuint64_t)CIN, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.N};
  const cuuint64_t strides[4] = {(cuuint64_t)p.x_x_stride * 2, (cuuint64_t)p.x_y_stride * 2, (cuuint64_t)p.x_d_stride * 2,
                                 (cuuint64_t)p.x_n_stride * 2};         // bytes, dims 1..4
  const cuuint32_t box[5] = {(cuuint32_t)CIN, (cuuint32_t)T::ROWV, (cuuint32_t)T::HH, (cuuint32_t)T::HD, 1u};   // ROWV = HW (+1 for Cin 8)
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<float*>(p.x), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CIN == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (CIN == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_NONE),
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <bool TMA>
static int launch_conv_s2_t(const bmv_conv3d_params& p, cudaStream_t st, const CUtensorMap& map) {
  using T = ConvS2;
  const size_t smem = (size_t)T::TILE_BYTES + (size_t)T::W_WORDS * 4 + 128;
  g_last_conv3d_tma = TMA ? 1 : 0;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_k3s2_c8_mma_kernel<TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3d_k3s2_c8_mma_kernel<TMA>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("bmv_conv3d_k3: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const int Do = (p.D - 1) / 2 + 1, Ho = (p.H - 1) / 2 + 1, Wo = (p.W - 1) / 2 + 1;
  const int64_t blocks = (int64_t)p.N * ((Do + T::TD - 1) / T::TD) * ((Ho + T::TH - 1) / T::TH) * ((Wo + T::TW - 1) / T::TW);
  conv3d_k3s2_c8_mma_kernel<TMA><<<(unsigned)blocks, kConvThreads, smem, st>>>(p, Do, Ho, Wo, map);
  return check_launch("bmv_conv3d_k3");
}

static int launch_conv_s2(const bmv_conv3d_params& p, cudaStream_t st) {
  using T = ConvS2;
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  EncodeTiledFn enc = encode_tiled_fn();
  if (p.in_half && !p.no_tma && enc && p.x_x_stride == 8) {
    const cuuint64_t dims[5] = {8, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.N};
    const cuuint64_t strides[4] = {(cuuint64_t)p.x_x_stride * 2, (cuuint64_t)p.x_y_stride * 2, (cuuint64_t)p.x_d_stride * 2,
                                   (cuuint64_t)p.x_n_stride * 2};
    const cuuint32_t box[5] = {8, (cuuint32_t)T::ROWV, (cuuint32_t)T::HH, (cuuint32_t)T::HD, 1u};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<float*>(p.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
      return launch_conv_s2_t<true>(p, st, map);
  }
  if (p.in_half) {
    set_error("bmv_conv3d_k3: fp16 input with stride 2 needs the TMA path (voxels contiguous along x, driver support)");
    return BMV_ERR_UNSUPPORTED_SHAPE;
  }
  return launch_conv_s2_t<false>(p, st, map);
}

template <int CIN, int NTILES, bool IN_HALF, bool TMA>
static int launch_conv_t(const bmv_conv3d_params& p, cudaStream_t st, const CUtensorMap& map) {
  using T = ConvTile<CIN, NTILES>;
  const size_t smem = (size_t)T::TILE_BYTES + (size_t)T::W_WORDS * 4 + 128;
  g_last_conv3d_tma = TMA ? 1 : 0;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_k3_mma_kernel<CIN, NTILES, IN_HALF, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3d_k3_mma_kernel<CIN, NTILES, IN_HALF, TMA>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("bmv_conv3d_k3: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const int64_t blocks = (int64_t)p.N * ((p.D + T::TD - 1) / T::TD) * ((p.H + T::TH - 1) / T::TH) * ((p.W + T::TW - 1) / T::TW);
  conv3d_k3_mma_kernel<CIN, NTILES, IN_HALF, TMA><<<(unsigned)blocks, kConvThreads, smem, st>>>(p, map);
  return check_launch("bmv_conv3d_k3");
}

template <int CIN, int NTILES>
static int launch_conv(const bmv_conv3d_params& p, cudaStream_t st) {
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  if constexpr (CIN == 16 || (CIN == 32 && NTILES == 1) || (CIN == 8 && NTILES == 2)) {
    if (p.in_half && !p.no_tma && make_input_map<CIN, NTILES>(p, &map)) return launch_conv_t<CIN, NTILES, true, true>(p, st, map);
  }
  return p.in_half ? launch_conv_t<CIN, NTILES, true, false>(p, st, map) : launch_conv_t<CIN, NTILES, false, false>(p, st, map);
}

}  // namespace bmv

extern "C" BMV_API int bmv_conv3d_k3(const bmv_conv3d_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_conv3d_k3");
  using namespace bmv;
  BMV_REQUIRE(p && p->x && p->wfrag && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3: null pointer");
  BMV_REQUIRE(p->N >= 1 && p->D >= 1 && p->H >= 1 && p->W >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3: bad size");
  {
    const int m = p->in_half ? 8 : 4;                                   // elements per 16 bytes
    BMV_REQUIRE(p->x_x_stride % m == 0 && p->x_y_stride % m == 0 && p->x_d_stride % m == 0 && p->x_n_stride % m == 0 &&
                    ((uintptr_t)p->x & 15) == 0 && ((uintptr_t)p->wfrag & 15) == 0,
                BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3: input must be channels-last with 16-byte aligned voxels");
  }
  BMV_REQUIRE(!p->out_half || (!p->out2 && p->Cout % 2 == 0 && p->o_x_stride % 2 == 0 && p->o_y_stride % 2 == 0 &&
                               p->o_d_stride % 2 == 0 && p->o_n_stride % 2 == 0 && ((uintptr_t)p->out & 3) == 0),
              BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3: fp16 output needs even Cout, a single output tensor and 4-byte aligned voxels");
  BMV_REQUIRE(!p->out2 || (p->split >= 1 && p->split < p->Cout), BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3: bad split");
  cudaStream_t st = (cudaStream_t)stream;
  BMV_REQUIRE(p->stride == 0 || p->stride == 1 || p->stride == 2, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3: stride must be 1 or 2");
  if (p->stride == 2) {
    BMV_REQUIRE(!p->in_scale, BMV_ERR_UNSUPPORTED_SHAPE, "bmv_conv3d_k3: in_scale is implemented by the stride-1 kernel only");
    BMV_REQUIRE(p->Cin == 8 && p->Cout % 2 == 0 && p->Cout <= 16 && !p->out2, BMV_ERR_UNSUPPORTED_SHAPE,
                "bmv_conv3d_k3: stride 2 is instantiated for Cin=8, even Cout<=16, single output (got Cin=%d, Cout=%d)", p->Cin, p->Cout);
    BMV_REQUIRE(p->o_x_stride % 2 == 0 && p->o_y_stride % 2 == 0 && p->o_d_stride % 2 == 0 && p->o_n_stride % 2 == 0 &&
                    ((uintptr_t)p->out & (p->out_half ? 3 : 7)) == 0, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3: stride-2 output alignment");
    return launch_conv_s2(*p, st);
  }
  if (p->Cin == 16 && p->Cout <= 8) return launch_conv<16, 1>(*p, st);
  if (p->Cin == 16 && p->Cout <= 16) return launch_conv<16, 2>(*p, st);
  if (p->Cin == 32 && p->Cout <= 8) return launch_conv<32, 1>(*p, st);
  if (p->Cin == 8 && p->Cout <= 16) return launch_conv<8, 2>(*p, st);
  set_error("bmv_conv3d_k3: (Cin=%d, Cout=%d) not instantiated (16->16, 32->8, 8->16)", p->Cin, p->Cout);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}

// 1 if the calling thread's last stride-1 bmv_conv3d_k3 launch staged its input tile with TMA, else 0
extern "C" BMV_API int bmv_conv3d_k3_last_used_tma(void) { return bmv::g_last_conv3d_tma; }

// words (uint32) of the fragment-ordered weight buffer for a (Cin, Cout) pair, -1 if not instantiated
extern "C" BMV_API int bmv_conv3d_k3_weight_words(int Cin, int Cout) {
  if (Cin == 16 && Cout <= 8) return bmv::ConvTile<16, 1>::W_WORDS;
  if (Cin == 16 && Cout <= 16) return bmv::ConvTile<16, 2>::W_WORDS;
  if (Cin == 32 && Cout <= 8) return bmv::ConvTile<32, 1>::W_WORDS;
  if (Cin == 8 && Cout <= 16) return bmv::ConvTile<8, 2>::W_WORDS;
  return -1;
}
