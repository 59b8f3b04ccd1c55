// Fused top-down step of the 2-D feature pyramid (a kept cuDNN module):
//     out = bilinear_up2x(prev, align_corners=True) + conv1x1(lateral_in) + bias
// i.e. `_upsample_add(x, lat(c))` of reference lib/networks/enerf/feature_net.py:24-33.  As three library
// calls (1x1 conv, upsample, add) this step moves the 32-channel full-resolution tensor through HBM
// three times and took 1.43 ms of the 4.9 ms FPN on B200 (profiles/round1_fpn_layers.md); fused it is one
// read of each input and one write.  Channels-last tensors; one thread per (pixel, 4 output channels).
#include "bmv_internal.cuh"

namespace bmv {

template <int CIN>
__global__ void __launch_bounds__(256) fpn_topdown_kernel(bmv_fpn_topdown_params p) {
  // weights transposed to [input][32 outputs]: the 8 channel groups of a warp read 128 contiguous bytes
  __shared__ __align__(16) float sW[32 * CIN + 32];
  for (int i = threadIdx.x; i < 32 * CIN; i += blockDim.x) sW[(i % CIN) * 32 + i / CIN] = p.weight[i];
  if (threadIdx.x < 32) sW[32 * CIN + threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;
  __syncthreads();
  // grid (x blocks, y, n): no integer division on the per-thread path
  const int gx = blockIdx.x * blockDim.x + threadIdx.x;
  if (gx >= p.W * 8) return;
  const int cg = gx & 7, x = gx >> 3, y = blockIdx.y, n = blockIdx.z;
  const int64_t pix = ((int64_t)n * p.H + y) * p.W + x;
  // 1x1 lateral conv for 4 output channels
  const float* in = p.lateral_in + pix * CIN;
  float v[CIN];
#pragma unroll
  for (int i = 0; i < CIN; i += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(in + i));
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
  float acc[4];
  {
    const float4 bb = *reinterpret_cast<const float4*>(sW + 32 * CIN + cg * 4);
    acc[0] = bb.x; acc[1] = bb.y; acc[2] = bb.z; acc[3] = bb.w;
#pragma unroll
    for (int i = 0; i < CIN; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(sW + i * 32 + cg * 4);
      acc[0] = fmaf(w.x, v[i], acc[0]); acc[1] = fmaf(w.y, v[i], acc[1]);
      acc[2] = fmaf(w.z, v[i], acc[2]); acc[3] = fmaf(w.w, v[i], acc[3]);
    }
  }
  // bilinear x2 upsample of prev (align_corners=True), ATen order of operations
  const int Hp = p.H / 2, Wp = p.W / 2;
  const UpCoord uy = up_coord(y, Hp, p.H), ux = up_coord(x, Wp, p.W);
  const float* pb = p.prev + (int64_t)n * Hp * Wp * 32 + cg * 4;
  const float4 a = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i0 * Wp + ux.i0) * 32));
  const float4 b = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i0 * Wp + ux.i1) * 32));
  const float4 c = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i1 * Wp + ux.i0) * 32));
  const float4 d = __ldg(reinterpret_cast<const float4*>(pb + ((int64_t)uy.i1 * Wp + ux.i1) * 32));
  float4 r;
  r.x = (uy.l0 * (ux.l0 * a.x + ux.l1 * b.x) + uy.l1 * (ux.l0 * c.x + ux.l1 * d.x)) + acc[0];
  r.y = (uy.l0 * (ux.l0 * a.y + ux.l1 * b.y) + uy.l1 * (ux.l0 * c.y + ux.l1 * d.y)) + acc[1];
  r.z = (uy.l0 * (ux.l0 * a.z + ux.l1 * b.z) + uy.l1 * (ux.l0 * c.z + ux.l1 * d.z)) + acc[2];
  r.w = (uy.l0 * (ux.l0 * a.w + ux.l1 * b.w) + uy.l1 * (ux.l0 * c.w + ux.l1 * d.w)) + acc[3];
  *reinterpret_cast<float4*>(p.out + pix * 32 + cg * 4) = r;
}

}  // namespace bmv

extern "C" BMV_API int bmv_fpn_topdown(const bmv_fpn_topdown_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_fpn_topdown");
  using namespace bmv;
  BMV_REQUIRE(p && p->prev && p->lateral_in && p->weight && p->out, BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_topdown: null pointer");
  BMV_REQUIRE(p->N >= 1 && p->H >= 2 && p->W >= 2 && p->H % 2 == 0 && p->W % 2 == 0, BMV_ERR_INVALID_ARGUMENT,
              "bmv_fpn_topdown: H, W must be even and >= 2");
  BMV_REQUIRE(((uintptr_t)p->prev & 15) == 0 && ((uintptr_t)p->lateral_in & 15) == 0 && ((uintptr_t)p->out & 15) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_topdown: tensors must be 16-byte aligned");
  BMV_REQUIRE(p->H <= 65535 && p->N <= 65535, BMV_ERR_INVALID_ARGUMENT, "bmv_fpn_topdown: H and N must be <= 65535");
  const dim3 blocks((unsigned)ceil_div64((int64_t)p->W * 8, 256), (unsigned)p->H, (unsigned)p->N);
  cudaStream_t st = (cudaStream_t)stream;
  if (p->Cin == 8) fpn_topdown_kernel<8><<<blocks, 256, 0, st>>>(*p);
  else if (p->Cin == 16) fpn_topdown_kernel<16><<<blocks, 256, 0, st>>>(*p);
  else {
    set_error("bmv_fpn_topdown: Cin=%d not instantiated (8, 16)", p->Cin);
    return BMV_ERR_UNSUPPORTED_SHAPE;
  }
  return check_launch("bmv_fpn_topdown");
}
