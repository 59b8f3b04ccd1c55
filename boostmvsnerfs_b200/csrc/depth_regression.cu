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
  const float* lg = p.logits + i;
  const float* pl = p.planes + (int64_t)i * p.planes_pix_stride;
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

}  // namespace bmv

extern "C" BMV_API int bmv_depth_regression(const bmv_depth_regression_params* p, bmv_stream_t stream) {
  using namespace bmv;
  BMV_REQUIRE(p && p->logits && p->planes && p->depth && p->std, BMV_ERR_INVALID_ARGUMENT,
              "bmv_depth_regression: null pointer");
  BMV_REQUIRE(p->D >= 1 && p->h >= 1 && p->w >= 1, BMV_ERR_INVALID_ARGUMENT, "bmv_depth_regression: bad size");
  const int hw = p->h * p->w;
  depth_regression_kernel<<<(hw + 255) / 256, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_depth_regression");
}
