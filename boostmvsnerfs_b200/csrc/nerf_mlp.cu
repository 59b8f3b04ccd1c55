// K5: stand-alone fused per-sample MLP (function-level boundary: NeRF.forward of the reference,
// lib/networks/enerf/nerf.py:29-43).  The per-frame path uses the same nerf_mlp_eval() inside the
// fused render kernel (render_fused.cu); this entry exists for op-level parity and for callers that
// already hold vox_feat / img_feat_rgb_dir tensors.
#include "nerf_mlp.cuh"

namespace bmv {

constexpr int kMlpThreads = 128;

template <int F, int V>
__global__ void __launch_bounds__(kMlpThreads, 3) nerf_mlp_kernel(bmv_nerf_mlp_params p) {
  using L = MlpLayout<F>;
  constexpr int ROW = V * (F + 4);
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;                          // packed weights
  float* sin = smem + L::TOTAL;              // [kMlpThreads][ROW] staged inputs
  for (int i = threadIdx.x * 4; i < L::TOTAL; i += kMlpThreads * 4)
    *reinterpret_cast<float4*>(sw + i) = __ldg(reinterpret_cast<const float4*>(p.weights + i));
  const int64_t s0 = (int64_t)blockIdx.x * kMlpThreads;
  const int n = (int)min((int64_t)kMlpThreads, p.P - s0);
  {  // coalesced copy of this block's img_feat rows (block start is 16-B aligned: 128*ROW*4 B)
    const float* src = p.img_feat + s0 * ROW;
    const int total = n * ROW;
    const int vec = total / 4;
    for (int i = threadIdx.x; i < vec; i += kMlpThreads)
      *reinterpret_cast<float4*>(sin + i * 4) = __ldg(reinterpret_cast<const float4*>(src) + i);
    for (int i = vec * 4 + threadIdx.x; i < total; i += kMlpThreads) sin[i] = __ldg(src + i);
  }
  __syncthreads();
  if ((int)threadIdx.x >= n) return;
  const int64_t s = s0 + threadIdx.x;
  float vox[8];
  {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.vox_feat + s * 8));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.vox_feat + s * 8) + 1);
    vox[0] = a.x; vox[1] = a.y; vox[2] = a.z; vox[3] = a.w; vox[4] = b.x; vox[5] = b.y; vox[6] = b.z; vox[7] = b.w;
  }
  float f[V][F + 4];
#pragma unroll
  for (int v = 0; v < V; ++v)
#pragma unroll
    for (int c = 0; c < F + 4; ++c) f[v][c] = sin[threadIdx.x * ROW + v * (F + 4) + c];
  const float4 o = nerf_mlp_eval<F, V>(sw, vox, f);
  reinterpret_cast<float4*>(p.raw)[s] = o;
}

template <int F, int V>
static int launch_mlp(const bmv_nerf_mlp_params& p, cudaStream_t st) {
  using L = MlpLayout<F>;
  const size_t smem = (size_t)(L::TOTAL + kMlpThreads * V * (F + 4)) * sizeof(float);
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(nerf_mlp_kernel<F, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("bmv_nerf_mlp: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  nerf_mlp_kernel<F, V><<<(unsigned)ceil_div64(p.P, kMlpThreads), kMlpThreads, smem, st>>>(p);
  return check_launch("bmv_nerf_mlp");
}

}  // namespace bmv

extern "C" BMV_API int bmv_nerf_mlp_weight_count(int feat_ch) {
  switch (feat_ch) {
    case 11: return bmv::MlpLayout<11>::TOTAL;
    case 35: return bmv::MlpLayout<35>::TOTAL;
    default: return -1;
  }
}

extern "C" BMV_API int bmv_nerf_mlp(const bmv_nerf_mlp_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_nerf_mlp");
  using namespace bmv;
  BMV_REQUIRE(p != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_nerf_mlp: null params");
  BMV_REQUIRE(p->P >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_nerf_mlp: negative sample count");
  if (p->P == 0) return BMV_OK;
  BMV_REQUIRE(p->vox_feat && p->img_feat && p->weights && p->raw, BMV_ERR_INVALID_ARGUMENT,
              "bmv_nerf_mlp: null device pointer");
  BMV_REQUIRE(((uintptr_t)p->weights & 15) == 0 && ((uintptr_t)p->vox_feat & 15) == 0 &&
                  ((uintptr_t)p->img_feat & 15) == 0 && ((uintptr_t)p->raw & 15) == 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_nerf_mlp: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->feat_ch == 11 && p->V == 3) return launch_mlp<11, 3>(*p, st);
  set_error("bmv_nerf_mlp: (feat_ch=%d, V=%d) not instantiated (available: 11x3)", p->feat_ch, p->V);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}
