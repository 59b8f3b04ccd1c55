// Tensor-core ConvTranspose3d(k=3, stride=2, padding=1, output_padding=1) + bias + skip connection:
//     out = skip + convT(x) + bias                      (x at half resolution, out/skip at full)
// for the decoder steps of the kept 3-D cost regularisers (`x = conv0 + self.conv11(x)`,
// `x = conv2 + self.conv9(x)`; reference lib/networks/enerf/cost_reg_net.py:23-31,40-44,62-70,80-82; BN folded).
//
// Why: on cuDNN the full-resolution step is a transposed-convolution kernel (0.36 ms at C2 level 1)
// followed by a separate elementwise add (3 more passes over the 134 MB tensor); fused, the skip
// tensor is read once and the result written once.
//
// A stride-2 transposed convolution is 8 ordinary sub-convolutions, one per output parity class
// (pz,py,px): per dimension an even output 2m takes tap k=1 of input m, an odd output 2m+1 takes
// tap k=2 of input m and tap k=0 of input m+1 — 1x, 2x, 2x, 4x, ... 8x taps, 27 in total.
// Implicit GEMM on mma.sync.m16n8k16: M = 16 consecutive INPUT x positions, K = 16 input channels,
// N = 8 output channels; a warp handles one (d, y, 16 x) input position block: 8 A fragments
// (2x2x2 input shifts) feed the 27 MMAs of its 8 output classes.  Operands fp16, accumulation fp32
// (TF32-class, same gating as conv3d_mma.cu).
#include "bmv_internal.cuh"
#include "conv_mma.cuh"

namespace bmv {

constexpr int kCtThreads = 256;
constexpr int kCtWarps = kCtThreads / 32;

template <int CIN, int COUT>
struct CtCfg {
  static constexpr int KT = CIN / 16, NT = COUT / 8;
  static constexpr int VS = CIN * 2;                                    // bytes per staged voxel
  static constexpr int TD = CIN == 32 ? 2 : 4, TH = CIN == 32 ? 2 : 4, TW = 32;   // input positions per CTA (32->16 runs at 1/4 res: small tiles = enough CTAs)
  static constexpr int HD = TD + 1, HH = TH + 1, HW = TW + 1;           // +1: odd outputs read input m+1
  static constexpr int ROWB = HW * VS;
  static constexpr int TILE_BYTES = HD * HH * ROWB;
  static constexpr int W_WORDS = 27 * KT * NT * 32 * 2;
  static constexpr bool BREG = false;
  static constexpr int JOBS = TD * TH * (TW / 16);
  __device__ static __forceinline__ int swz(int v) { return CIN == 16 ? (v >> 2) & 1 : (v >> 1) & 3; }
};

template <int CIN, int COUT>
__global__ void __launch_bounds__(kCtThreads, (CIN == 16 ? 4 : 2)) convT3d_k3s2_mma_kernel(bmv_convT3d_params p) {
  using C = CtCfg<CIN, COUT>;
  constexpr int KT = C::KT, NT = C::NT;
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* tile = smem;
  const uint2* wfrag = reinterpret_cast<const uint2*>(smem + C::TILE_BYTES);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.wfrag);
    uint4* dst = reinterpret_cast<uint4*>(smem + C::TILE_BYTES);
    for (int i = threadIdx.x; i < C::W_WORDS / 4; i += kCtThreads) dst[i] = __ldg(src + i);
  }
  const int tiles_w = (p.W + C::TW - 1) / C::TW, tiles_h = (p.H + C::TH - 1) / C::TH, tiles_d = (p.D + C::TD - 1) / C::TD;
  int b = blockIdx.x;
  const int tw = b % tiles_w; b /= tiles_w;
  const int th = b % tiles_h; b /= tiles_h;
  const int td = b % tiles_d; b /= tiles_d;
  const int n = b;
  const int x0 = tw * C::TW, y0 = th * C::TH, d0 = td * C::TD;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // ---- stage the input tile (+1 halo on the high side) as fp16, one warp per row (see conv3d_mma.cu)
  if (p.in_half) {                                                      // fp16 input: straight 16-byte copies
    const __half* xin = reinterpret_cast<const __half*>(p.x) + (int64_t)n * p.x_n_stride;
    constexpr int CH8 = CIN / 8;
    constexpr int PER_ROW = C::HW * CH8;
    constexpr int ROWS = C::HD * C::HH;
    constexpr int P = (PER_ROW + 31) / 32;
    constexpr int VPP = 32 / CH8;
    const int c8 = lane % CH8, hx0 = lane / CH8;
    const int64_t lane_off = (int64_t)(x0 + hx0) * p.x_x_stride + c8 * 8, pass_off = (int64_t)VPP * p.x_x_stride;
    for (int r = warp; r < ROWS; r += kCtWarps) {
      const int hd = r / C::HH, hy = r - hd * C::HH;
      const int gy = y0 + hy, gd = d0 + hd;
      const bool row_ok = gy < p.H && gd < p.D;
      const __half* src = xin + (int64_t)gd * p.x_d_stride + (int64_t)gy * p.x_y_stride + lane_off;
      uint4 val[P];
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int hx = hx0 + VPP * k;
        val[k] = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok && hx < C::HW && x0 + hx < p.W) val[k] = __ldg(reinterpret_cast<const uint4*>(src + k * pass_off));
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int hx = hx0 + VPP * k;
        if (hx < C::HW) *reinterpret_cast<uint4*>(tile + r * C::ROWB + hx * C::VS + ((c8 ^ C::swz(hx)) << 4)) = val[k];
      }
    }
  } else
  {
    const float* xin = p.x + (int64_t)n * p.x_n_stride;
    constexpr int CH4 = CIN / 4;
    constexpr int PER_ROW = C::HW * CH4;
    constexpr int ROWS = C::HD * C::HH;
    constexpr int P = (PER_ROW + 31) / 32;
    constexpr int VPP = 32 / CH4;
    const int c4 = lane % CH4, hx0 = lane / CH4;
    const int64_t lane_off = (int64_t)(x0 + hx0) * p.x_x_stride + c4 * 4, pass_off = (int64_t)VPP * p.x_x_stride;
    for (int r = warp; r < ROWS; r += kCtWarps) {
      const int hd = r / C::HH, hy = r - hd * C::HH;
      const int gy = y0 + hy, gd = d0 + hd;
      const bool row_ok = gy < p.H && gd < p.D;
      const float* src = xin + (int64_t)gd * p.x_d_stride + (int64_t)gy * p.x_y_stride + lane_off;
      float4 val[P];
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int hx = hx0 + VPP * k;
        val[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok && hx < C::HW && x0 + hx < p.W) val[k] = __ldg(reinterpret_cast<const float4*>(src + k * pass_off));
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int hx = hx0 + VPP * k;
        if (hx < C::HW)
          *reinterpret_cast<uint2*>(tile + r * C::ROWB + hx * C::VS + (((c4 >> 1) ^ C::swz(hx)) << 4) + (c4 & 1) * 8) = pack_half4(val[k]);
      }
    }
  }
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(tile);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;
  uint2 breg[C::BREG ? 27 * KT * NT : 1];
  if (C::BREG) {
#pragma unroll
    for (int i = 0; i < 27 * KT * NT; ++i) breg[i] = wfrag[i * 32 + lane];
  }
  float bias0[NT], bias1[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    bias0[nt] = p.bias ? __ldg(p.bias + nt * 8 + 2 * t) : 0.f;
    bias1[nt] = p.bias ? __ldg(p.bias + nt * 8 + 2 * t + 1) : 0.f;
  }
  const float* skip = p.skip ? p.skip + (int64_t)n * p.s_n_stride : nullptr;
  float* out = p.out + (int64_t)n * p.o_n_stride;

  for (int job = warp; job < C::JOBS; job += kCtWarps) {
    const int mx = job % (C::TW / 16), my = (job / (C::TW / 16)) % C::TH, md = job / ((C::TW / 16) * C::TH);
    if (d0 + md >= p.D || y0 + my >= p.H || x0 + mx * 16 >= p.W) continue;          // warp-uniform
    float acc[8][NT][4];                                                // [class = pz*4+py*2+px]
#pragma unroll
    for (int c = 0; c < 8; ++c)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        acc[c][nt][0] = bias0[nt]; acc[c][nt][1] = bias1[nt]; acc[c][nt][2] = bias0[nt]; acc[c][nt][3] = bias1[nt];
      }
#pragma unroll
    for (int sz = 0; sz < 2; ++sz)
#pragma unroll
      for (int sy = 0; sy < 2; ++sy)
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
          const int v = mx * 16 + lrow + sx;
          const uint32_t rowaddr = tile_s + ((md + sz) * C::HH + (my + sy)) * C::ROWB + v * C::VS;
#pragma unroll
          for (int kt = 0; kt < KT; ++kt) {
            uint32_t a[4];
            ldmatrix_x4(a, rowaddr + (((kt * 2 + lhi) ^ C::swz(v)) << 4));
            // shift s in a dimension serves (parity 0, tap 1) and (parity 1, tap 2) when s = 0, (parity 1, tap 0) when s = 1
#pragma unroll
            for (int oz = 0; oz < 2 - sz; ++oz)
#pragma unroll
              for (int oy = 0; oy < 2 - sy; ++oy)
#pragma unroll
                for (int ox = 0; ox < 2 - sx; ++ox) {
                  const int pz = sz ? 1 : oz, kz = sz ? 0 : 1 + oz;
                  const int py = sy ? 1 : oy, ky = sy ? 0 : 1 + oy;
                  const int px = sx ? 1 : ox, kx = sx ? 0 : 1 + ox;
                  const int tap = (kz * 3 + ky) * 3 + kx, cls = pz * 4 + py * 2 + px;
#pragma unroll
                  for (int nt = 0; nt < NT; ++nt) {
                    const int wi = (tap * KT + kt) * NT + nt;
                    const uint2 bw = C::BREG ? breg[C::BREG ? wi : 0] : wfrag[wi * 32 + lane];
                    hmma16816(acc[cls][nt], a, bw.x, bw.y);
                  }
                }
          }
        }
    // ---- epilogue: + skip, store.  Lane holds input positions m0+g and m0+g+8, channels 2t, 2t+1 of each n-tile.
    const int m_lo = x0 + mx * 16 + g, m_hi = m_lo + 8;
#pragma unroll
    for (int cls = 0; cls < 8; ++cls) {
      const int pz = cls >> 2, py = (cls >> 1) & 1, px = cls & 1;
      const int od = 2 * (d0 + md) + pz, oy = 2 * (y0 + my) + py;
      const int64_t rowo = (int64_t)od * p.o_d_stride + (int64_t)oy * p.o_y_stride;
      const int64_t rows = (int64_t)od * p.s_d_stride + (int64_t)oy * p.s_y_stride;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = nt * 8 + 2 * t;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int m = h ? m_hi : m_lo;
          if (m >= p.W) continue;
          const int ox = 2 * m + px;
          float2 v = make_float2(acc[cls][nt][2 * h], acc[cls][nt][2 * h + 1]);
          if (skip) {
            if (p.skip_half) {
              const __half2 hs = *reinterpret_cast<const __half2*>(reinterpret_cast<const __half*>(p.skip) + (int64_t)n * p.s_n_stride + rows +
                                                                   (int64_t)ox * p.s_x_stride + c);
              const float2 s = __half22float2(hs);
              v.x += s.x; v.y += s.y;
            } else {
              const float2 s = __ldg(reinterpret_cast<const float2*>(skip + rows + (int64_t)ox * p.s_x_stride + c));
              v.x += s.x; v.y += s.y;
            }
          }
          if (p.out_half) {
            *reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(p.out) + (int64_t)n * p.o_n_stride + rowo + (int64_t)ox * p.o_x_stride + c) = pack_half2_sat(v.x, v.y);
          } else {
            *reinterpret_cast<float2*>(out + rowo + (int64_t)ox * p.o_x_stride + c) = v;
          }
        }
      }
    }
  }
}

template <int CIN, int COUT>
static int launch_convT(const bmv_convT3d_params& p, cudaStream_t st) {
  using C = CtCfg<CIN, COUT>;
  const size_t smem = (size_t)C::TILE_BYTES + (size_t)C::W_WORDS * 4;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(convT3d_k3s2_mma_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(convT3d_k3s2_mma_kernel<CIN, COUT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) {
      set_error("bmv_convT3d_k3s2: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const int64_t blocks = (int64_t)p.N * ((p.D + C::TD - 1) / C::TD) * ((p.H + C::TH - 1) / C::TH) * ((p.W + C::TW - 1) / C::TW);
  convT3d_k3s2_mma_kernel<CIN, COUT><<<(unsigned)blocks, kCtThreads, smem, st>>>(p);
  return check_launch("bmv_convT3d_k3s2");
}

}  // namespace bmv

extern "C" BMV_API int bmv_convT3d_k3s2(const bmv_convT3d_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_convT3d_k3s2");
  using namespace bmv;
  BMV_REQUIRE(p && p->x && p->wfrag && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_convT3d_k3s2: null pointer");
  BMV_REQUIRE(p->N >= 1 && p->D >= 1 && p->H >= 1 && p->W >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_convT3d_k3s2: bad size");
  const int xm = p->in_half ? 8 : 4;
  BMV_REQUIRE(p->x_x_stride % xm == 0 && p->x_y_stride % xm == 0 && p->x_d_stride % xm == 0 && p->x_n_stride % xm == 0 &&
                  ((uintptr_t)p->x & 15) == 0 && ((uintptr_t)p->wfrag & 15) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_convT3d_k3s2: input must be channels-last with 16-byte aligned voxels");
  BMV_REQUIRE(p->o_x_stride % 2 == 0 && p->o_y_stride % 2 == 0 && p->o_d_stride % 2 == 0 && p->o_n_stride % 2 == 0 &&
                  ((uintptr_t)p->out & (p->out_half ? 3 : 7)) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_convT3d_k3s2: output must be channels-last with 8-byte aligned voxels");
  BMV_REQUIRE(!p->skip || (p->s_x_stride % 2 == 0 && p->s_y_stride % 2 == 0 && p->s_d_stride % 2 == 0 &&
                           p->s_n_stride % 2 == 0 && ((uintptr_t)p->skip & (p->skip_half ? 3 : 7)) == 0),
              BMV_ERR_INVALID_ARGUMENT, "bmv_convT3d_k3s2: skip must be channels-last with 8-byte aligned voxels");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->Cin == 16 && p->Cout == 8) return launch_convT<16, 8>(*p, st);
  if (p->Cin == 32 && p->Cout == 16) return launch_convT<32, 16>(*p, st);
  set_error("bmv_convT3d_k3s2: (Cin=%d, Cout=%d) not instantiated (16->8, 32->16)", p->Cin, p->Cout);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}

extern "C" BMV_API int bmv_convT3d_k3s2_weight_words(int Cin, int Cout) {
  if (Cin == 16 && Cout == 8) return bmv::CtCfg<16, 8>::W_WORDS;
  if (Cin == 32 && Cout == 16) return bmv::CtCfg<32, 16>::W_WORDS;
  return -1;
}
