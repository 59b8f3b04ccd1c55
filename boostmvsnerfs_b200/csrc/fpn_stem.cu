// Fused stem of the 2-D feature pyramid: conv0 = ConvBnReLU(3,8,3) -> ConvBnReLU(8,8,3) at full
// resolution (reference lib/networks/enerf/feature_net.py:7-9,29; BN folded into weight + bias).
//
// Why: on cuDNN the two layers take 0.26 + 0.20 ms for the 6 source views of C2 while moving only
// 37 MB in and 100 MB out (plus the 100 MB intermediate, written and read); fused, the intermediate
// lives in shared memory and the step is bounded by the output write.
//
// Stage A: image tile (+2 halo) -> shared memory (fp32).
// Stage B: first convolution on CUDA cores in fp32 (K = 27 is too ragged for an MMA tile), ReLU, stored
//          as fp16 with a 1-pixel halo (zero outside the image = the second convolution's padding).
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
  __shared__ __align__(16) float s_img[kStIY * kStIX * 3];
  __shared__ __align__(16) unsigned char s_mid[kStMY * kStMRowB];
  __shared__ __align__(16) float s_w0[27 * 8 + 8];                       // [tap*3+c][8 outputs], bias
  __shared__ __align__(16) uint2 s_w1[3 * 2 * 32];
  for (int i = threadIdx.x; i < 216; i += kStThreads) {
    const int o = i / 27, r = i - o * 27, c = r / 9, tap = r - c * 9;    // weight (8,3,3,3) row-major: o, c, ky, kx
    s_w0[(tap * 3 + c) * 8 + o] = __ldg(p.w0 + i);
  }
  if (threadIdx.x < 8) s_w0[216 + threadIdx.x] = p.b0 ? __ldg(p.b0 + threadIdx.x) : 0.f;
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
  // ---- stage B: one thread per intermediate pixel, all 8 channels
  for (int i = threadIdx.x; i < kStMY * (kStMX + 1); i += kStThreads) {
    const int my = i / (kStMX + 1), mx = i - my * (kStMX + 1);
    const int y = y0 + my - 1, x = x0 + mx - 1;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (mx < kStMX && y >= 0 && y < p.H && x >= 0 && x < p.W) {
      lo = *reinterpret_cast<const float4*>(s_w0 + 216);
      hi = *reinterpret_cast<const float4*>(s_w0 + 220);
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* px = s_img + ((my + ky) * kStIX + (mx + kx)) * 3;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float v = px[c];
            const float4 wl = *reinterpret_cast<const float4*>(s_w0 + ((ky * 3 + kx) * 3 + c) * 8);
            const float4 wh = *reinterpret_cast<const float4*>(s_w0 + ((ky * 3 + kx) * 3 + c) * 8 + 4);
            lo.x = fmaf(wl.x, v, lo.x); lo.y = fmaf(wl.y, v, lo.y); lo.z = fmaf(wl.z, v, lo.z); lo.w = fmaf(wl.w, v, lo.w);
            hi.x = fmaf(wh.x, v, hi.x); hi.y = fmaf(wh.y, v, hi.y); hi.z = fmaf(wh.z, v, hi.z); hi.w = fmaf(wh.w, v, hi.w);
          }
        }
      lo.x = fmaxf(lo.x, 0.f); lo.y = fmaxf(lo.y, 0.f); lo.z = fmaxf(lo.z, 0.f); lo.w = fmaxf(lo.w, 0.f);
      hi.x = fmaxf(hi.x, 0.f); hi.y = fmaxf(hi.y, 0.f); hi.z = fmaxf(hi.z, 0.f); hi.w = fmaxf(hi.w, 0.f);
    }
    const uint2 a = pack_half4(lo), b = pack_half4(hi);
    *reinterpret_cast<uint4*>(s_mid + my * kStMRowB + mx * 16) = make_uint4(a.x, a.y, b.x, b.y);
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
