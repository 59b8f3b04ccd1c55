// Fused stem of the 2-D feature pyramid: conv0 = ConvBnReLU(3,8,3) -> ConvBnReLU(8,8,3) at full
// resolution (reference lib/networks/enerf/feature_net.py:7-9,29; BN folded into weight + bias).
//
// Why: on cuDNN the two layers take 0.26 + 0.20 ms for the 6 source views of C2 while moving only
// 37 MB in and 100 MB out (plus the 100 MB intermediate, written and read); fused, the intermediate
// lives in shared memory and the step is bounded by the output write.
//
// Stage A: image tile (+2 halo) -> shared memory (fp32).
// Stage B: first convolution on tensor cores as well (round 2; on CUDA cores it was ~300 instructions per pixel — 27
//          taps x (one pixel load, two weight-vector loads, 8 FMAs) — and the kernel sat at 81 % of the L1 / shared-memory
//          pipe): M = 16 consecutive pixels, K = 32 = the 27 (tap, colour) products + zero padding, N = 8, fp32 bias
//          in the accumulators; every lane gathers its A-fragment entries straight from the fp32 image tile (8 fixed
//          offsets), ReLU, stored as fp16 with a 1-pixel halo (zero outside the image = the second convolution's padding).
// Stage C: second convolution on tensor cores (mma.sync.m16n8k16, 8 input channels: two horizontal taps
//          per MMA as in conv3d_mma.cu), bias + ReLU, fp32 (or fp16: out_half) channels-last output.
// fp16 operands in stage C make the result TF32-class: the host uses this kernel only when
// torch.backends.cudnn.allow_tf32 is set (inference_plan.py).
#include "bmv_internal.cuh"
#include "conv_mma.cuh"

namespace bmv {

constexpr int kStThreads = 256;
constexpr int kStTY = 8, kStTX = 64, kStWY = 4;
constexpr int kStIY = kStTY + 4, kStIX = kStTX + 4;                      // image tile
constexpr int kStMY = kStTY + 2, kStMX = kStTX + 2;                      // intermediate tile
constexpr int kStMRowB = (kStMX + 1) * 16;                               // +1 pixel: the unpaired tap reads x+3

__global__ void __launch_bounds__(kStThreads, 3) fpn_stem_kernel(bmv_fpn_stem_params p) {
  __shared__ __align__(16) float s_img[kStIY * kStIX * 3 + 2];            // + {1.0, 0.0}: constant slots for the padding columns of stage B
  __shared__ __align__(16) unsigned char s_mid[kStMY * kStMRowB];
  __shared__ __align__(16) uint2 s_w1[3 * 2 * 32];
  if (threadIdx.x == 0) { s_img[kStIY * kStIX * 3] = 1.f; s_img[kStIY * kStIX * 3 + 1] = 0.f; }
  if (threadIdx.x < 3 * 2 * 32) s_w1[threadIdx.x] = __ldg(reinterpret_cast<const uint2*>(p.wfrag1) + threadIdx.x);
  const int tiles_x = (p.W + kStTX - 1) / kStTX;
  const int x0 = (blockIdx.x % tiles_x) * kStTX, y0 = (blockIdx.x / tiles_x) * kStTY, n = blockIdx.y;
  // ---- stage A
  {
    const float* img = p.x + (int64_t)n * p.x_n_stride;
    for (int i = threadIdx.x; i < kStIY * kStIX; i += kStThreads) {
      const int iy = i / kStIX, ix = i - iy * kStIX;
      const int y = y0 + iy - 2, x = x0 + ix - 2;
      float r = 0.f, g = 0.f, b = 0.f;
      if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
        const float* px = img + (int64_t)y * p.x_y_stride + (int64_t)x * p.x_x_stride;
        r = __ldg(px); g = __ldg(px + p.x_c_stride); b = __ldg(px + 2 * p.x_c_stride);
      }
      s_img[i * 3] = r; s_img[i * 3 + 1] = g; s_img[i * 3 + 2] = b;
      // optional by-product: the image as (N,H,W,4) channels-last for the colour fetch of the render kernels
      if (p.rgb4 && iy >= 2 && iy < 2 + kStTY && ix >= 2 && ix < 2 + kStTX && y < p.H && x < p.W)
        *reinterpret_cast<float4*>(p.rgb4 + (((int64_t)n * p.H + y) * p.W + x) * 4) = make_float4(r, g, b, 0.f);
    }
  }
  __syncthreads();
  // ---- stage B: 3 -> 8 convolution on tensor cores; a warp takes 16 consecutive intermediate pixels (row-major over the
  // kStMY x kStMX tile).  K index k < 27 is (tap = k / 3, colour = k % 3), k >= 27 zero.
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    // B fragments (weights (8,3,3,3) row-major: o, c, ky, kx) and the lane's 8 A-fragment sources: element offset in
    // s_img relative to the pixel's window origin (mul = 1), or the absolute slot of the constant 1 / 0 (mul = 0)
    uint32_t bw[2][2];
    int koff[2][2][2], kmul[2][2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        float wv[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = ks * 16 + r * 8 + 2 * t + e;
          const int tap = k / 3, c = k - tap * 3, ky = tap / 3, kx = tap - ky * 3;
          wv[e] = k < 27 ? __ldg(p.w0 + g * 27 + c * 9 + tap) : 0.f;
          koff[ks][r][e] = k < 27 ? (ky * kStIX + kx) * 3 + c : kStIY * kStIX * 3 + 1;
          kmul[ks][r][e] = k < 27 ? 1 : 0;
        }
        bw[ks][r] = pack_half2_sat(wv[0], wv[1]);
      }
    const float bias_a = p.b0 ? __ldg(p.b0 + 2 * t) : 0.f, bias_b = p.b0 ? __ldg(p.b0 + 2 * t + 1) : 0.f;
    constexpr int NPIX = kStMY * kStMX, NSEG = (NPIX + 15) / 16;
    for (int seg = warp; seg < NSEG; seg += kStThreads / 32) {
      int base3[2], myv[2], mxv[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int pi = min(seg * 16 + g + 8 * h, NPIX - 1);
        myv[h] = pi / kStMX; mxv[h] = pi - myv[h] * kStMX;
        base3[h] = (myv[h] * kStIX + mxv[h]) * 3;
      }
      float acc[4] = {bias_a, bias_b, bias_a, bias_b};                  // fp32 bias, as cuDNN adds it
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t a[4];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float v0 = s_img[base3[h] * kmul[ks][r][0] + koff[ks][r][0]], v1 = s_img[base3[h] * kmul[ks][r][1] + koff[ks][r][1]];
            a[r * 2 + h] = pack_half2_sat(v0, v1);
          }
        hmma16816(acc, a, bw[ks][0], bw[ks][1]);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (seg * 16 + g + 8 * h >= NPIX) continue;
        const int y = y0 + myv[h] - 1, x = x0 + mxv[h] - 1;
        const bool ok = (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
        const uint32_t pk = ok ? pack_half2_sat(fmaxf(acc[2 * h], 0.f), fmaxf(acc[2 * h + 1], 0.f)) : 0u;
        *reinterpret_cast<uint32_t*>(s_mid + myv[h] * kStMRowB + mxv[h] * 16 + t * 4) = pk;
      }
    }
    // the extra column the unpaired tap of stage C reads (x + 3): zeros
    if (threadIdx.x < kStMY) *reinterpret_cast<uint4*>(s_mid + threadIdx.x * kStMRowB + kStMX * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  // ---- stage C: 8 -> 8 convolution on tensor cores; a warp owns kStWY rows x 16 pixels
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t mid_s = (uint32_t)__cvta_generic_to_shared(s_mid);
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lhi = lane >> 4;
  uint2 breg[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) breg[i] = s_w1[i * 32 + lane];
  const float b0 = p.b1 ? __ldg(p.b1 + 2 * t) : 0.f, b1 = p.b1 ? __ldg(p.b1 + 2 * t + 1) : 0.f;
  float* out = p.out + (int64_t)n * p.H * p.W * 8;
  constexpr int JOBS = (kStTY / kStWY) * (kStTX / 16);
  for (int job = warp; job < JOBS; job += kStThreads / 32) {
    const int jx = job % (kStTX / 16), yb = (job / (kStTX / 16)) * kStWY;
    if (y0 + yb >= p.H || x0 + jx * 16 >= p.W) continue;
    float acc[kStWY][4];
#pragma unroll
    for (int oy = 0; oy < kStWY; ++oy) { acc[oy][0] = b0; acc[oy][1] = b1; acc[oy][2] = b0; acc[oy][3] = b1; }
    const uint32_t lane_base = mid_s + (jx * 16 + lrow + lhi) * 16;       // k 8..15 = the next pixel
#pragma unroll
    for (int py = 0; py < kStWY + 2; ++py) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t a[4];
        ldmatrix_x4(a, lane_base + (yb + py) * kStMRowB + j * 32);
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const int oy = py - dy;
          if (oy < 0 || oy >= kStWY) continue;
          hmma16816(acc[oy], a, breg[dy * 2 + j].x, breg[dy * 2 + j].y);
        }
      }
    }
    const int gx0 = x0 + jx * 16 + g, gx1 = gx0 + 8;
#pragma unroll
    for (int oy = 0; oy < kStWY; ++oy) {
      const int gy = y0 + yb + oy;
      if (gy >= p.H) continue;
      const float2 r0 = make_float2(fmaxf(acc[oy][0], 0.f), fmaxf(acc[oy][1], 0.f));
      const float2 r1 = make_float2(fmaxf(acc[oy][2], 0.f), fmaxf(acc[oy][3], 0.f));
      if (p.out_half) {             // 4 bytes per lane; the 8 pixels of a warp-level store are 128 contiguous bytes
        __half* o16 = reinterpret_cast<__half*>(p.out) + (int64_t)n * p.H * p.W * 8;
        if (gx0 < p.W) *reinterpret_cast<uint32_t*>(o16 + ((int64_t)gy * p.W + gx0) * 8 + 2 * t) = pack_half2_sat(r0.x, r0.y);
        if (gx1 < p.W) *reinterpret_cast<uint32_t*>(o16 + ((int64_t)gy * p.W + gx1) * 8 + 2 * t) = pack_half2_sat(r1.x, r1.y);
        continue;
      }
      if (gx0 < p.W) *reinterpret_cast<float2*>(out + ((int64_t)gy * p.W + gx0) * 8 + 2 * t) = r0;
      if (gx1 < p.W) *reinterpret_cast<float2*>(out + ((int64_t)gy * p.W + gx1) * 8 + 2 * t) = r1;
      if (p.out_s2d) {                                                  // (N, H/2, W/2, [py][px][8])
        float* zrow = p.out_s2d + (((int64_t)n * (p.H >> 1) + (gy >> 1)) * (p.W >> 1)) * 32 + (gy & 1) * 16 + 2 * t;
        if (gx0 < p.W) *reinterpret_cast<float2*>(zrow + (int64_t)(gx0 >> 1) * 32 + (gx0 & 1) * 8) = r0;
        if (gx1 < p.W) *reinterpret_cast<float2*>(zrow + (int64_t)(gx1 >> 1) * 32 + (gx1 & 1) * 8) = r1;
      }
    }
  }
}

}  // namespace bmv

extern "C" BMV_API int bmv_fpn_stem(const bmv_fpn_stem_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_fpn_stem");
  using namespace bmv;
  BMV_REQUIRE(p && p->x && p->w0 && p->wfrag1 && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_stem: null pointer");
  BMV_REQUIRE(p->N >= 1 && p->N <= 65535 && p->H >= 1 && p->W >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_stem: bad size");
  BMV_REQUIRE(((uintptr_t)p->out & 7) == 0 && ((uintptr_t)p->wfrag1 & 7) == 0 && ((uintptr_t)p->rgb4 & 15) == 0 &&
                  ((uintptr_t)p->out_s2d & 7) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_stem: out / out_s2d / wfrag1 must be 8-byte, rgb4 16-byte aligned");
  BMV_REQUIRE(!p->out_half || !p->out_s2d, BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_stem: the fp16 output has no space-to-depth copy");
  BMV_REQUIRE(!p->out_s2d || (p->H % 2 == 0 && p->W % 2 == 0), BMV_ERR_INVALID_ARGUMENT,
              "bmv_fpn_stem: the space-to-depth output needs even H and W");
  const dim3 grid((unsigned)(((p->W + kStTX - 1) / kStTX) * ((p->H + kStTY - 1) / kStTY)), (unsigned)p->N);
  fpn_stem_kernel<<<grid, kStThreads, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_fpn_stem");
}
