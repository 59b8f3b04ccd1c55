// 3x3x3 / stride-1 / pad-1 convolution (+bias, ReLU, split outputs) on the 5th-generation tensor cores:
// TMA -> shared memory -> tcgen05.mma -> tensor memory -> registers -> HBM.  Same contract and the same
// fp16-operand / fp32-accumulate arithmetic as conv3d_mma.cu (TF32-class, gated by the host the same way).
//
// Why: ncu on the TMA-staged mma.sync kernels (profiles/round1j) shows them bound by the legacy HMMA path
// (pipe 52-59 % busy, math-pipe-throttle stalls) at 0.25-0.4 of the HBM roofline.  With tcgen05 the implicit
// GEMM needs no fragments in registers at all: M = 128 consecutive x voxels of one output row, N = 16 output
// channels (Cout padded), K = 16 input values per instruction, and the A operand of tap (dz,dy,dx) is the
// staged halo tile itself, addressed by a shared-memory descriptor whose start is shifted by the tap:
//     Cin  8: voxels are 16 bytes = one K-chunk; rows (voxels) 16 B apart = the SWIZZLE_NONE core-matrix pitch,
//             the second K-chunk is the NEXT voxel (LBO = 16): taps dx, dx+1 in one MMA, 2 K-steps per (dz,dy)
//             (the partner of dx = 2 has zero weights);
//     Cin 16: voxels are 32 bytes = the SWIZZLE_32B row the TMA unit writes (absolute-address XOR, the same on
//             the tensor-core side), one K-step per tap.
// An output row's accumulator is 16 TMEM columns (lane = voxel), the rows of an output plane are adjacent column
// groups.  With N = 16 an MMA is bound by reading its 4 KB A operand from shared memory, not by its 8 clk of math
// (ncu on the final kernel: tensor pipe 85 % busy, tensor-core shared-memory wavefronts 71 % of the L1 data pipe),
// so ONE MMA serves every output row an input row contributes to: input row hy of plane z' feeds
// output rows oy = hy-2 .. hy (dy = hy - oy) of plane z'-dz, whose accumulators are adjacent -> N = 16..48 with
// B = a row range of the stacked weights [W(dz,2); W(dz,1); W(dz,0)].  That is 3 (TH+2) KS / TH MMAs per output
// row instead of 9 KS.  Accumulators are zeroed by one MMA against a zero B operand (an MMA spans rows that were
// and were not written before, so the accumulate flag cannot do it).  The epilogue thread of voxel x reads its
// 16 columns with one tcgen05.ld, adds the bias, applies ReLU and writes the voxel's channels as 16/32-byte stores.
// The kernel is persistent and warp-specialised (TMA producer warp, one MMA-issuer warp per output plane, eight
// epilogue warps; NSLOT staged tiles + accumulator sets in flight, mbarriers / tcgen05.commit between the roles) — see
// the comment above the kernel.  Version history with the measurements that drove it: profiles/round1l_tcgen05.md.
#include <cuda.h>

#include <cstring>
#include <cuda_fp16.h>

#include "bmv_internal.cuh"
#include "umma.cuh"

namespace bmv {

template <int CIN> struct UConv;
template <> struct UConv<8> {
  // The (C, W) dimensions are merged into TMA box rows of 128 voxels = 256 8-byte elements (the box limit): 36 rows per
  // tile instead of 4716 16-byte ones (measured: 100 -> 96 us — the copy was not the bottleneck, but the rows are also
  // a power of two).  An MMA still spans 128 voxels, of which the last two read past the row: 126 valid outputs per x tile.
  static constexpr int VS = 16, KS = 2, TH = 4, TD = 4, ROWV = 128, TWV = 126, NSLOT = 2;
  static constexpr uint32_t LAYOUT = 0, A_LBO = 16, A_SBO = 128;
  __device__ static __forceinline__ uint32_t step_off(int j) { return (uint32_t)j * 32u; }      // voxels 2j, 2j+1
};
template <> struct UConv<16> {
  static constexpr int VS = 32, KS = 3, TH = 4, TD = 2, ROWV = 130, TWV = 128, NSLOT = 2;
  static constexpr uint32_t LAYOUT = 6, A_LBO = 16, A_SBO = 256;                                 // SWIZZLE_32B: 8 rows x 32 B
  __device__ static __forceinline__ uint32_t step_off(int j) { return (uint32_t)j * 32u; }      // voxel j
};
template <int CIN> struct UTile {
  using C = UConv<CIN>;
  static constexpr int TW = C::TWV, TH = C::TH, TD = C::TD, HH = TH + 2, HD = TD + 2;   // TW: valid outputs per x tile
  static constexpr int ROWV = C::ROWV, ROWB = ROWV * C::VS;
  static constexpr int TILE_BYTES = HD * HH * ROWB;
  static constexpr int WS_BYTES = 2 * 48 * 16;               // stacked [W(dz,2); W(dz,1); W(dz,0)] of one (dz, k-step): (K/8, 48, 8) fp16
  static constexpr int W_BYTES = 3 * C::KS * WS_BYTES;
  static constexpr int Z_BYTES = 2 * 16 * TH * 16;           // zero B operand (N = 16 TH) that clears a plane's accumulators
  static constexpr int ROWS = TD * TH, NSLOT = C::NSLOT;      // NSLOT tiles in flight (staged tile + accumulators each)
  static constexpr uint32_t TMEM_COLS = ROWS * 16 <= 32 ? 32 : (ROWS * 16 <= 64 ? 64 : (ROWS * 16 <= 128 ? 128 : 256));
  static constexpr int TILE_PAD = 64;                                          // zeros behind the last row (read by the MMA rows past it)
  static constexpr int TILE_STRIDE = (TILE_BYTES + TILE_PAD + 1023) / 1024 * 1024;   // staged tiles, each 1024-byte aligned
  static constexpr size_t SMEM = (size_t)NSLOT * TILE_STRIDE + W_BYTES + Z_BYTES + 1024;
};

__device__ __forceinline__ void umma_mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void umma_tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

__device__ __forceinline__ void umma_tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
               ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void umma_mbar_arrive(uint32_t mbar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar) : "memory");
}

// Epilogue of one CTA's tiles for the calling thread (voxel = lane of its TMEM quarter).  MODE 0: fp32 features
// (8 channels, 32-byte voxels) + channel 8 to out2 (the merged heads); MODE 1: fp16, 8 channels (16-byte voxels);
// MODE 2: fp32, 8 or 16 channels, single tensor; MODE 3: any strides / channel counts (per-element stores).
template <int CIN, int MODE>
__device__ __forceinline__ void uconv_epilogue(const bmv_conv3d_params& p, uint64_t (*s_acc_full)[UTile<CIN>::TD], uint64_t* s_acc_empty,
                                               uint32_t tmem_base, int n_tiles, int tiles_w, int tiles_h, int tiles_d) {
  using T = UTile<CIN>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = warp & 3, vx = q * 32 + lane;
  const int eset = (warp - (1 + T::TD)) >> 2;                // which half of the output planes this warp drains
  const int split = p.out2 ? p.split : p.Cout;
  const bool has_bias = p.bias != nullptr, relu = p.relu != 0;
  float bias[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) bias[c] = (has_bias && c < p.Cout) ? __ldg(p.bias + c) : 0.f;
  const int64_t o_n = p.o_n_stride, o_d = p.o_d_stride, o_y = p.o_y_stride, o_x = p.o_x_stride;
  const int64_t o2_n = p.o2_n_stride, o2_d = p.o2_d_stride, o2_y = p.o2_y_stride, o2_x = p.o2_x_stride;
  const int D = p.D, H = p.H, W = p.W;
  int it = 0;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
    const int slot = it % T::NSLOT, ph = (it / T::NSLOT) & 1;
    int b = t;
    const int tw = b % tiles_w; b /= tiles_w;
    const int th = b % tiles_h; b /= tiles_h;
    const int td = b % tiles_d; b /= tiles_d;
    const int n = b, x0 = tw * T::TW, y0 = th * T::TH, d0 = td * T::TD;
    const int gx = x0 + vx;
    const bool store = gx < W && vx < T::TW;
    const int rows_y = min(T::TH, H - y0), rows_d = min(T::TD, D - d0);
    const uint32_t trow = tmem_base + (uint32_t)(slot * T::TMEM_COLS) + ((uint32_t)(q * 32) << 16);
    const int64_t vo = (int64_t)n * o_n + (int64_t)d0 * o_d + (int64_t)y0 * o_y + (int64_t)gx * o_x;
    const int64_t vo2 = (int64_t)n * o2_n + (int64_t)d0 * o2_d + (int64_t)y0 * o2_y + (int64_t)gx * o2_x;
    for (int od = eset; od < T::TD; od += 2) {
      mbar_wait(smem_u32(&s_acc_full[slot][od]), ph);
      __syncwarp();
      tc_fence_after();
      if (od >= rows_d) continue;                           // uniform
      for (int oy = 0; oy < rows_y; ++oy) {
        float v[16];
        tmem_ld16(trow + (uint32_t)((od * T::TH + oy) * 16), v);
        if (!store) continue;
        if (has_bias) {
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] += bias[c];
        }
        if (relu) {
#pragma unroll
          for (int c = 0; c < 16; ++c) v[c] = fmaxf(v[c], 0.f);
        }
        const int64_t ro = vo + od * o_d + oy * o_y;
        if (MODE == 0) {
          float* o = p.out + ro;
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
          p.out2[vo2 + od * o2_d + oy * o2_y] = v[8];
        } else if (MODE == 1) {
          __half* o = reinterpret_cast<__half*>(p.out) + ro;
          uint4 h;
          h.x = pack_half2_sat(v[0], v[1]);
          h.y = pack_half2_sat(v[2], v[3]);
          h.z = pack_half2_sat(v[4], v[5]);
          h.w = pack_half2_sat(v[6], v[7]);
          *reinterpret_cast<uint4*>(o) = h;
        } else if (MODE == 2) {
          float* o = p.out + ro;
          *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
          if (split == 16) {
            *reinterpret_cast<float4*>(o + 8) = make_float4(v[8], v[9], v[10], v[11]);
            *reinterpret_cast<float4*>(o + 12) = make_float4(v[12], v[13], v[14], v[15]);
          }
        } else if (p.out_half) {
          __half* o = reinterpret_cast<__half*>(p.out) + ro;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (2 * c < p.Cout) *reinterpret_cast<uint32_t*>(o + 2 * c) = pack_half2_sat(v[2 * c], v[2 * c + 1]);
        } else {
          float* o = p.out + ro;
#pragma unroll
          for (int c = 0; c < 16; ++c)
            if (c < split) o[c] = v[c];
          if (p.out2) {
            float* o2 = p.out2 + vo2 + od * o2_d + oy * o2_y;
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c >= split && c < p.Cout) o2[c - split] = v[c];
          }
        }
      }
    }
    tc_fence_before();                                      // this thread's tcgen05.ld of the slot are complete (wait::ld inside tmem_ld16)
    umma_mbar_arrive(smem_u32(&s_acc_empty[slot]));
  }
}

// warp 0: TMA producer; warps 1..TD: MMA issuers (one output plane each); then 8 epilogue warps, two per TMEM lane
// quarter (the first four take the even output planes, the other four the odd ones)
template <int CIN> struct UcThreads { static constexpr int value = (1 + UTile<CIN>::TD + 8) * 32; };

// Persistent, warp-specialised: every CTA (one per SM) walks tiles blockIdx.x, +gridDim.x, ... through a NSLOT-slot
// pipeline.  A slot = one staged halo tile + one set of accumulators (half of the CTA's TMEM columns):
//   producer : wait empty[slot] -> one 5-D TMA copy of the next tile -> full[slot]
//   issuers  : wait full[slot] and acc_empty[slot] -> MMAs -> commit acc_full[slot][plane] per plane, empty[slot] at the end
//   epilogue : wait acc_full[slot][plane] -> tcgen05.ld, bias, ReLU, stores -> arrive acc_empty[slot]
// so the copy of tile i+1 and the epilogue of tile i-1 run under the MMAs of tile i; TMEM allocation, barrier set-up
// and the weight copy happen once per CTA.
template <int CIN>
__global__ void __launch_bounds__(UcThreads<CIN>::value, 1) conv3d_k3_umma_kernel(bmv_conv3d_params p, const __grid_constant__ CUtensorMap tmap,
                                                                     int n_tiles) {
  using T = UTile<CIN>;
  using C = UConv<CIN>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // swizzled TMA boxes and the tensor core XOR absolute address bits: keep the tiles 1024-byte aligned
  unsigned char* tile0 = smem_raw + ((1024u - ((uint32_t)__cvta_generic_to_shared(smem_raw) & 1023u)) & 1023u);
  unsigned char* wsm = tile0 + T::NSLOT * T::TILE_STRIDE;
  __shared__ __align__(8) uint64_t s_full[T::NSLOT], s_empty[T::NSLOT], s_acc_empty[T::NSLOT], s_acc_full[T::NSLOT][T::TD];
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tiles_w = (p.W + T::TW - 1) / T::TW, tiles_h = (p.H + T::TH - 1) / T::TH, tiles_d = (p.D + T::TD - 1) / T::TD;
  if (tid == 0) {
    for (int s2 = 0; s2 < T::NSLOT; ++s2) {
      mbar_init(smem_u32(&s_full[s2]), 1);
      mbar_init(smem_u32(&s_empty[s2]), T::TD);             // every issuer warp commits
      mbar_init(smem_u32(&s_acc_empty[s2]), 256);           // every epilogue thread arrives
      for (int i = 0; i < T::TD; ++i) mbar_init(smem_u32(&s_acc_full[s2][i]), 1);
    }
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), T::NSLOT * T::TMEM_COLS);
  {                                                         // weights (already in operand order) + the zero operand, once per CTA
    const uint4* src = reinterpret_cast<const uint4*>(p.wfrag);
    uint4* dst = reinterpret_cast<uint4*>(wsm);
    for (int i = tid; i < T::W_BYTES / 16; i += UcThreads<CIN>::value) dst[i] = __ldg(src + i);
    for (int i = tid; i < T::Z_BYTES / 16; i += UcThreads<CIN>::value) dst[T::W_BYTES / 16 + i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid < T::NSLOT * T::TILE_PAD / 16)
      *reinterpret_cast<uint4*>(tile0 + (tid / (T::TILE_PAD / 16)) * T::TILE_STRIDE + T::TILE_BYTES + (tid % (T::TILE_PAD / 16)) * 16) =
          make_uint4(0u, 0u, 0u, 0u);
  }
  proxy_fence_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
        const int slot = it % T::NSLOT, ph = (it / T::NSLOT) & 1;
        int b = t;
        const int tw = b % tiles_w; b /= tiles_w;
        const int th = b % tiles_h; b /= tiles_h;
        const int td = b % tiles_d; b /= tiles_d;
        mbar_wait(smem_u32(&s_empty[slot]), ph ^ 1);        // the MMAs that read this slot two tiles ago are done
        umma_mbar_expect_tx(smem_u32(&s_full[slot]), (uint32_t)T::TILE_BYTES);
        if (CIN == 8)                                       // merged (C, W) rows of 8-byte elements: 2 per voxel
          umma_tma_load_4d(smem_u32(tile0 + slot * T::TILE_STRIDE), &tmap, smem_u32(&s_full[slot]), (tw * T::TW - 1) * 2, th * T::TH - 1,
                           td * T::TD - 1, b);
        else
          umma_tma_load_5d(smem_u32(tile0 + slot * T::TILE_STRIDE), &tmap, smem_u32(&s_full[slot]), 0, tw * T::TW - 1, th * T::TH - 1,
                           td * T::TD - 1, b);
      }
    }
  } else if (warp <= T::TD) {
    // ------------------------------------------------------------------ MMA issuers: warp 1 + od owns output plane od.
    // The whole warp runs the loop on warp-uniform values (the plane index comes from a shuffle so that the compiler
    // keeps descriptors in uniform registers) and one elected lane issues: with `if (lane == 0)` around everything each
    // MMA cost ~125 clk of ELECT / R2UR.BROADCAST / branch sequences and two issuers were the bottleneck of the kernel.
    constexpr uint32_t A_HI = (C::A_SBO >> 4) | (1u << 14) | (C::LAYOUT << 29);
    constexpr uint32_t B_HI = (128u >> 4) | (1u << 14);
    const uint32_t w_lo = ((smem_u32(wsm) & 0x3FFFFu) >> 4) | ((768u >> 4) << 16);                  // stacked weights: LBO = 48 rows x 16 B
    const uint32_t z_lo = (((smem_u32(wsm) + T::W_BYTES) & 0x3FFFFu) >> 4) | (((16u * T::TH * 16u) >> 4) << 16);
    const int od = __shfl_sync(0xffffffffu, warp, 0) - 1;
    const bool issue = elect_one();
    int it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
      const int slot = it % T::NSLOT, ph = (it / T::NSLOT) & 1;
      mbar_wait(smem_u32(&s_full[slot]), ph);
      mbar_wait(smem_u32(&s_acc_empty[slot]), ph ^ 1);
      tc_fence_after();
      const uint32_t a_lo = (((smem_u32(tile0 + slot * T::TILE_STRIDE) + (uint32_t)(od * T::HH * T::ROWB)) & 0x3FFFFu) >> 4) |
                            ((C::A_LBO >> 4) << 16);      // input plane od of the slot (dz = 0)
      const uint32_t dplane = tmem_base + (uint32_t)(slot * T::TMEM_COLS + od * T::TH * 16);
      if (issue) {
        umma_f16_lohi<false>(dplane, a_lo + (uint32_t)(((T::HH + 1) * T::ROWB) >> 4), A_HI, z_lo, B_HI, umma_idesc(16 * T::TH));
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
          for (int hy = 0; hy < T::HH; ++hy) {
            constexpr int TH = T::TH;
            const int oy_min = hy - 2 > 0 ? hy - 2 : 0, oy_max = hy < TH - 1 ? hy : TH - 1;
            const int cnt = oy_max - oy_min + 1, b0 = 2 - (hy - oy_min);
#pragma unroll
            for (int j = 0; j < C::KS; ++j) {
              const uint32_t bl = w_lo + (uint32_t)((((dz * C::KS + j) * T::WS_BYTES) + b0 * 256) >> 4);
              const uint32_t al = a_lo + (uint32_t)((((dz * T::HH + hy) * T::ROWB) + (int)C::step_off(j)) >> 4);
              umma_f16_lohi<true>(dplane + (uint32_t)(oy_min * 16), al, A_HI, bl, B_HI, umma_idesc(16 * cnt));
            }
          }
        umma_commit(smem_u32(&s_acc_full[slot][od]));
        umma_commit(smem_u32(&s_empty[slot]));              // this issuer's reads of the staged tile are complete
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------------ epilogue: thread = voxel x0 + 32 (warp % 4) + lane
    // (the first version spent 170 instructions per output row on 64-bit address arithmetic and per-row layout
    // decisions and was THE bottleneck of the kernel — four warps, one per scheduler; the store layout is now picked
    // once per launch and the row pointers advance by precomputed strides)
    const int mode = p.out_half ? ((p.Cout == 8 && p.o_x_stride % 8 == 0 && p.o_y_stride % 8 == 0 && p.o_d_stride % 8 == 0 &&
                                    p.o_n_stride % 8 == 0 && ((uintptr_t)p.out & 15) == 0) ? 1 : 3)
                                : ((p.out2 && p.split == 8 && p.Cout == 9 && p.o_x_stride % 4 == 0 && p.o_y_stride % 4 == 0 &&
                                    p.o_d_stride % 4 == 0 && p.o_n_stride % 4 == 0 && ((uintptr_t)p.out & 15) == 0) ? 0 :
                                   (!p.out2 && (p.Cout == 8 || p.Cout == 16) && p.o_x_stride % 4 == 0 && p.o_y_stride % 4 == 0 &&
                                    p.o_d_stride % 4 == 0 && p.o_n_stride % 4 == 0 && ((uintptr_t)p.out & 15) == 0) ? 2 : 3);
    if (mode == 0) uconv_epilogue<CIN, 0>(p, s_acc_full, s_acc_empty, tmem_base, n_tiles, tiles_w, tiles_h, tiles_d);
    else if (mode == 1) uconv_epilogue<CIN, 1>(p, s_acc_full, s_acc_empty, tmem_base, n_tiles, tiles_w, tiles_h, tiles_d);
    else if (mode == 2) uconv_epilogue<CIN, 2>(p, s_acc_full, s_acc_empty, tmem_base, n_tiles, tiles_w, tiles_h, tiles_d);
    else uconv_epilogue<CIN, 3>(p, s_acc_full, s_acc_empty, tmem_base, n_tiles, tiles_w, tiles_h, tiles_d);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, T::NSLOT * T::TMEM_COLS);
}

typedef CUresult (*UEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static UEncodeTiledFn u_encode_tiled_fn() {
  static UEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<UEncodeTiledFn>(f);
  }
  return fn;
}

template <int CIN>
static int launch_uconv(const bmv_conv3d_params& p, cudaStream_t st) {
  using T = UTile<CIN>;
  UEncodeTiledFn enc = u_encode_tiled_fn();
  BMV_REQUIRE(enc != nullptr, BMV_ERR_CUDA_LAUNCH, "bmv_conv3d_k3_umma: cuTensorMapEncodeTiled is not available");
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r;
  if (CIN == 8) {
    // (2W, H, D, N) of 8-byte elements (a voxel = 16 bytes = 2 elements), box = 128 voxels x HH x HD
    const cuuint64_t dims[4] = {(cuuint64_t)p.W * 2, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.N};
    const cuuint64_t strides[3] = {(cuuint64_t)p.x_y_stride * 2, (cuuint64_t)p.x_d_stride * 2, (cuuint64_t)p.x_n_stride * 2};
    const cuuint32_t box[4] = {(cuuint32_t)T::ROWV * 2, (cuuint32_t)T::HH, (cuuint32_t)T::HD, 1u};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 4, const_cast<float*>(p.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[5] = {(cuuint64_t)CIN, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.D, (cuuint64_t)p.N};
    const cuuint64_t strides[4] = {(cuuint64_t)p.x_x_stride * 2, (cuuint64_t)p.x_y_stride * 2, (cuuint64_t)p.x_d_stride * 2,
                                   (cuuint64_t)p.x_n_stride * 2};
    const cuuint32_t box[5] = {(cuuint32_t)CIN, (cuuint32_t)T::ROWV, (cuuint32_t)T::HH, (cuuint32_t)T::HD, 1u};
    r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<float*>(p.x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  BMV_REQUIRE(r == CUDA_SUCCESS, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3_umma: cuTensorMapEncodeTiled failed (%d)", (int)r);
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(conv3d_k3_umma_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv3d_k3_umma_kernel<CIN>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("bmv_conv3d_k3_umma: cannot reserve %zu B shared memory: %s", (size_t)T::SMEM, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const int64_t tiles = (int64_t)p.N * ((p.D + T::TD - 1) / T::TD) * ((p.H + T::TH - 1) / T::TH) * ((p.W + T::TW - 1) / T::TW);
  BMV_REQUIRE(tiles < (1ll << 31), BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3_umma: too many tiles");
  const unsigned blocks = (unsigned)(tiles < kNumSMs ? tiles : kNumSMs);    // persistent: one CTA per SM (it owns all TMEM it needs)
  conv3d_k3_umma_kernel<CIN><<<blocks, UcThreads<CIN>::value, T::SMEM, st>>>(p, map, (int)tiles);
  return check_launch("bmv_conv3d_k3_umma");
}

}  // namespace bmv

extern "C" BMV_API int bmv_conv3d_k3_umma(const bmv_conv3d_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_conv3d_k3_umma");
  using namespace bmv;
  BMV_REQUIRE(!p || !p->in_scale, BMV_ERR_UNSUPPORTED_SHAPE, "bmv_conv3d_k3_umma: in_scale is not implemented here");
  BMV_REQUIRE(p && p->x && p->wfrag && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3_umma: null pointer");
  BMV_REQUIRE(p->N >= 1 && p->D >= 1 && p->H >= 1 && p->W >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3_umma: bad size");
  BMV_REQUIRE(p->in_half && (p->stride == 0 || p->stride == 1) && p->x_x_stride == p->Cin, BMV_ERR_UNSUPPORTED_SHAPE,
              "bmv_conv3d_k3_umma: needs an fp16 channels-last input with contiguous voxels along x and stride 1");
  BMV_REQUIRE(p->x_y_stride % 8 == 0 && p->x_d_stride % 8 == 0 && p->x_n_stride % 8 == 0 && ((uintptr_t)p->x & 15) == 0 &&
                  ((uintptr_t)p->wfrag & 15) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3_umma: input rows / weights must be 16-byte aligned");
  BMV_REQUIRE(p->Cout >= 1 && p->Cout <= 16, BMV_ERR_UNSUPPORTED_SHAPE, "bmv_conv3d_k3_umma: Cout must be <= 16 (got %d)", p->Cout);
  BMV_REQUIRE(!p->out_half || (!p->out2 && p->Cout % 2 == 0 && p->o_x_stride % 2 == 0 && p->o_y_stride % 2 == 0 &&
                               p->o_d_stride % 2 == 0 && p->o_n_stride % 2 == 0 && ((uintptr_t)p->out & 3) == 0),
              BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3_umma: fp16 output needs even Cout, a single output tensor and 4-byte aligned voxels");
  BMV_REQUIRE(!p->out2 || (p->split >= 1 && p->split < p->Cout), BMV_ERR_INVALID_ARGUMENT, "bmv_conv3d_k3_umma: bad split");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->Cin == 8) return launch_uconv<8>(*p, st);
  if (p->Cin == 16) return launch_uconv<16>(*p, st);
  set_error("bmv_conv3d_k3_umma: Cin=%d not instantiated (8, 16)", p->Cin);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}

// bytes/4 of the operand-ordered weight buffer (mlp_pack.pack_conv3d_k3_umma), -1 if not instantiated
extern "C" BMV_API int bmv_conv3d_k3_umma_weight_words(int Cin, int Cout) {
  if (Cout < 1 || Cout > 16) return -1;
  if (Cin == 8) return bmv::UTile<8>::W_BYTES / 4;
  if (Cin == 16) return bmv::UTile<16>::W_BYTES / 4;
  return -1;
}
