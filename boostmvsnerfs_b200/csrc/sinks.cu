// Output sinks of a rendered frame, so that what leaves the GPU is what the reference's
// evaluator / visualiser actually consume instead of the fp32 frame:
//   * bmv_frame_psnr_accumulate: masked (and optionally centre-cropped) sum of squared errors + element count of a
//     predicted frame against the ground truth (reference lib/evaluators/enerf.py:45-71: pred/gt reshaped to
//     (h,w,3), `eval_center` crop of 10 %, `gt[mask]` vs `pred[mask]` into skimage's PSNR = 10 log10(1 / mse),
//     mse accumulated in float64);
//   * bmv_frame_to_u8: rgb -> uint8 with numpy's `(x * 255).astype(uint8)` truncation, depth min / max, and
//     depth -> uint8 `(d - min) / (max - min) * 255` (reference lib/visualizers/enerf.py:21-37).
// Both are single-pass HBM streams (16 / 28 bytes per pixel); the D2H copy shrinks from 16 to 4 bytes per pixel.
#include "bmv_internal.cuh"

namespace bmv {

__global__ void __launch_bounds__(256) frame_psnr_kernel(bmv_frame_psnr_params p) {
  double sse = 0.0;
  unsigned long long cnt = 0;
  const int64_t n = (int64_t)p.H * p.W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / p.W), x = (int)(i % p.W);
    if (y < p.crop_h || y >= p.H - p.crop_h || x < p.crop_w || x >= p.W - p.crop_w) continue;
    if (p.mask && !(__ldg(p.mask + i) >= 1)) continue;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double d = (double)__ldg(p.gt + i * 3 + c) - (double)__ldg(p.pred + i * 3 + c);
      sse = fma(d, d, sse);
    }
    cnt += 3;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sse += __shfl_xor_sync(0xffffffffu, sse, o);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  __shared__ double s_sse[8];
  __shared__ unsigned long long s_cnt[8];
  if ((threadIdx.x & 31) == 0) { s_sse[threadIdx.x >> 5] = sse; s_cnt[threadIdx.x >> 5] = cnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { sse += s_sse[i]; cnt += s_cnt[i]; }
    atomicAdd(p.sse, sse);
    atomicAdd(p.count, cnt);
  }
}

// order-preserving map float -> uint32 for atomicMin / atomicMax on floats of either sign
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void minmax_init_kernel(unsigned* s) {
  s[0] = 0xff800000u;   // ord(+inf): running minimum
  s[1] = 0x007fffffu;   // ord(-inf): running maximum
}

__global__ void __launch_bounds__(256) frame_rgb_u8_minmax_kernel(bmv_frame_to_u8_params p) {
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.R; i += (int64_t)gridDim.x * blockDim.x) {
    if (p.rgb_u8) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        // numpy: (pred * 255).astype(np.uint8) — fp32 multiply, truncation toward zero; values are convex
        // combinations of source colours in [0, 1], the clamp only guards the undefined out-of-range cast
        const float v = mul_rn(__ldg(p.rgb + i * 3 + c), 255.f);
        p.rgb_u8[i * 3 + c] = (uint8_t)(int)fminf(fmaxf(v, 0.f), 255.f);
      }
    }
    if (p.depth) {
      const float d = __ldg(p.depth + i);
      mn = fminf(mn, d);
      mx = fmaxf(mx, d);
    }
  }
  if (!p.depth) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(p.minmax_ord, f2ord(mn));
    atomicMax(p.minmax_ord + 1, f2ord(mx));
  }
}

__global__ void __launch_bounds__(256) frame_depth_u8_kernel(bmv_frame_to_u8_params p) {
  const float mn = ord2f(p.minmax_ord[0]), mx = ord2f(p.minmax_ord[1]);
  const float span = sub_rn(mx, mn);
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.minmax) { p.minmax[0] = mn; p.minmax[1] = mx; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.R; i += (int64_t)gridDim.x * blockDim.x) {
    // ((depth - depth.min()) / (depth.max() - depth.min()) * 255).astype(np.uint8): three separately rounded fp32 ops
    const float v = mul_rn(div_rn(sub_rn(__ldg(p.depth + i), mn), span), 255.f);
    p.depth_u8[i] = (uint8_t)(int)fminf(fmaxf(v, 0.f), 255.f);
  }
}

}  // namespace bmv

extern "C" BMV_API int bmv_frame_psnr_accumulate(const bmv_frame_psnr_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_frame_psnr_accumulate");
  using namespace bmv;
  BMV_REQUIRE(p && p->pred && p->gt && p->sse && p->count, BMV_ERR_INVALID_ARGUMENT, "bmv_frame_psnr_accumulate: null pointer");
  BMV_REQUIRE(p->H >= 1 && p->W >= 1 && p->crop_h >= 0 && p->crop_w >= 0 && 2 * p->crop_h < p->H && 2 * p->crop_w < p->W,
              BMV_ERR_INVALID_ARGUMENT, "bmv_frame_psnr_accumulate: bad size / crop");
  const int64_t want = ceil_div64((int64_t)p->H * p->W, 256 * 4);
  const unsigned blocks = (unsigned)(want < 4 * kNumSMs ? want : 4 * kNumSMs);
  frame_psnr_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_frame_psnr_accumulate");
}

extern "C" BMV_API int bmv_frame_to_u8(const bmv_frame_to_u8_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_frame_to_u8");
  using namespace bmv;
  BMV_REQUIRE(p && p->R >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_frame_to_u8: null params / empty frame");
  BMV_REQUIRE((p->rgb && p->rgb_u8) || (p->depth && p->depth_u8), BMV_ERR_INVALID_ARGUMENT, "bmv_frame_to_u8: nothing to convert");
  BMV_REQUIRE(!p->rgb_u8 || p->rgb, BMV_ERR_INVALID_ARGUMENT, "bmv_frame_to_u8: rgb_u8 without rgb");
  BMV_REQUIRE(!p->depth || (p->depth_u8 && p->minmax_ord), BMV_ERR_INVALID_ARGUMENT,
              "bmv_frame_to_u8: depth needs depth_u8 and the 2-word minmax_ord scratch");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t want = ceil_div64(p->R, 256 * 4);
  const unsigned blocks = (unsigned)(want < 4 * kNumSMs ? want : 4 * kNumSMs);
  if (p->depth) minmax_init_kernel<<<1, 1, 0, st>>>(p->minmax_ord);   // a kernel, not a copy: graph-capturable
  frame_rgb_u8_minmax_kernel<<<blocks, 256, 0, st>>>(*p);
  if (p->depth) frame_depth_u8_kernel<<<blocks, 256, 0, st>>>(*p);
  return check_launch("bmv_frame_to_u8");
}
