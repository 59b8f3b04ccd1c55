// K2: softmax over the D depth hypotheses of every pixel, expected depth and its standard
// deviation.  Reference semantics: lib/networks/enerf/utils.py:722-727 (depth_regression).
#include "bmv_internal.cuh"

namespace bmv {

// One thread per pixel; the D logits of a pixel are D coalesced plane reads (x fastest).
// Three passes over D from L1/L2 (max, sum-exp + mean, variance); D is 8..128 so the re-reads hit
// L1 — DRAM sees each input once.
__global__ void __launch_bounds__(256) depth_regression_kernel(bmv_depth_regression_params p) {
  const int hw = p.h * p.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const int b = blockIdx.y;                                             // chain
  const float* lg = p.logits + b * p.logits_b_stride + i;
  const float* pl = p.planes + b * p.planes_b_stride + (int64_t)i * p.planes_pix_stride;
  p.depth += (int64_t)b * hw; p.std += (int64_t)b * hw;
  float m = -INFINITY;
  for (int d = 0; d < p.D; ++d) m = fmaxf(m, __ldg(lg + (int64_t)d * hw));
  float den = 0.f;
  for (int d = 0; d < p.D; ++d) den += expf(__ldg(lg + (int64_t)d * hw) - m);
  float mean = 0.f;
  for (int d = 0; d < p.D; ++d) {
    float pr = div_rn(expf(__ldg(lg + (int64_t)d * hw) - m), den);
    float v = __ldg(pl + (int64_t)d * p.planes_d_stride);
    if (p.depth_inv) v = div_rn(1.f, fmaxf(v, 1e-6f));
    mean = add_rn(mean, mul_rn(pr, v));
  }
  float var = 0.f;
  for (int d = 0; d < p.D; ++d) {
    float pr = div_rn(expf(__ldg(lg + (int64_t)d * hw) - m), den);
    float v = __ldg(pl + (int64_t)d * p.planes_d_stride);
    if (p.depth_inv) v = div_rn(1.f, fmaxf(v, 1e-6f));
    float dv = sub_rn(v, mean);
    var = add_rn(var, mul_rn(pr, mul_rn(dv, dv)));
  }
  p.depth[i] = mean;
  p.std[i] = sqrtf(fmaxf(var, 1e-10f));
}

// Register-resident variant for the plane counts the cascade actually uses: all D logits and
// hypotheses of the pixel are fetched with D independent coalesced loads in flight (one DRAM
// latency instead of 3*D dependent L1 round trips), then reduced in registers.
template <int D>
__global__ void __launch_bounds__(128) depth_regression_reg_kernel(bmv_depth_regression_params p) {
  const int hw = p.h * p.w;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hw) return;
  const int b = blockIdx.y;                                             // chain
  const float* lg = p.logits + b * p.logits_b_stride + i;
  const float* pl = p.planes + b * p.planes_b_stride + (int64_t)i * p.planes_pix_stride;
  p.depth += (int64_t)b * hw; p.std += (int64_t)b * hw;
  float e[D], v[D];
#pragma unroll
  for (int d = 0; d < D; ++d) e[d] = __ldg(lg + (int64_t)d * hw);
#pragma unroll
  for (int d = 0; d < D; ++d) v[d] = __ldg(pl + (int64_t)d * p.planes_d_stride);
  float m = -INFINITY;
#pragma unroll
  for (int d = 0; d < D; ++d) m = fmaxf(m, e[d]);
  float den = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) { e[d] = expf(e[d] - m); den += e[d]; }
  float mean = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    e[d] = div_rn(e[d], den);
    if (p.depth_inv) v[d] = div_rn(1.f, fmaxf(v[d], 1e-6f));
    mean = add_rn(mean, mul_rn(e[d], v[d]));
  }
  float var = 0.f;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const float dv = sub_rn(v[d], mean);
    var = add_rn(var, mul_rn(e[d], mul_rn(dv, dv)));
  }
  p.depth[i] = mean;
  p.std[i] = sqrtf(fmaxf(var, 1e-10f));
}

}  // namespace bmv

extern "C" BMV_API int bmv_depth_regression(const bmv_depth_regression_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_depth_regression");
  using namespace bmv;
  BMV_REQUIRE(p && p->logits && p->planes && p->depth && p->std, BMV_ERR_INVALID_ARGUMENT,
              "bmv_depth_regression: null pointer");
  BMV_REQUIRE(p->D >= 1 && p->h >= 1 && p->w >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_depth_regression: bad size");
  BMV_REQUIRE(p->batch >= 0 && p->batch <= 65535, BMV_ERR_INVALID_ARGUMENT, "bmv_depth_regression: bad batch");
  const int hw = p->h * p->w;
  const unsigned nb = p->batch > 1 ? (unsigned)p->batch : 1u;
  cudaStream_t st = (cudaStream_t)stream;
  const dim3 g128((hw + 127) / 128, nb), g256((hw + 255) / 256, nb);
  switch (p->D) {
    case 8: depth_regression_reg_kernel<8><<<g128, 128, 0, st>>>(*p); break;
    case 16: depth_regression_reg_kernel<16><<<g128, 128, 0, st>>>(*p); break;
    case 32: depth_regression_reg_kernel<32><<<g128, 128, 0, st>>>(*p); break;
    case 64: depth_regression_reg_kernel<64><<<g128, 128, 0, st>>>(*p); break;
    default: depth_regression_kernel<<<g256, 256, 0, st>>>(*p);
  }
  return check_launch("bmv_depth_regression");
}
