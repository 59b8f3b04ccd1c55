// K3+K5 fused: depth-guided ray generation, sampling, feature gather, 3-D visibility and the
// per-sample MLP in ONE kernel; the 221 MB/chain of fetched features (vox_feat + img_feat_rgb_dir,
// SURVEY.md §11) never touch HBM.  Outputs per chain: raw (R,S,4), z_vals (R,S), visibility (R,S),
// which K4 (composite.cu) blends over the K chains.
// Reference call sequence replaced: render_rays of lib/networks/boost_enerf/network.py:123-149.
#include "nerf_mlp.cuh"
#include "raygen_common.cuh"

namespace bmv {

constexpr int kRenderThreads = 128;

template <int CF, int V>
__global__ void __launch_bounds__(kRenderThreads, 3) render_rays_kernel(bmv_render_rays_params rp) {
  constexpr int F = CF + 3;
  using L = MlpLayout<F>;
  const bmv_raygen_fetch_params& p = rp.g;
  extern __shared__ __align__(16) float smem[];
  float* sw = smem;
  __shared__ ViewCam cams[V];
  __shared__ float s_tar_c[3];
  __shared__ int s_view[V];
  for (int i = threadIdx.x * 4; i < L::TOTAL; i += kRenderThreads * 4)
    *reinterpret_cast<float4*>(sw + i) = __ldg(reinterpret_cast<const float4*>(rp.mlp_weights + i));
  if (threadIdx.x < V) s_view[threadIdx.x] = p.view[threadIdx.x];
  __syncthreads();
  for (int v = 0; v < V; ++v) load_cam(&cams[v], p.src_exts, p.src_ixts, p.src_centers, s_view[v], threadIdx.x);
  if (threadIdx.x < 3) s_tar_c[threadIdx.x] = p.tar_center[threadIdx.x];
  __syncthreads();
  const int64_t li = (int64_t)blockIdx.x * kRenderThreads + threadIdx.x;
  if (li >= p.n_rays) return;
  const RaySetup r = ray_setup(p, li);
  const float un = div_rn(r.fx, (float)(p.W - 1)), vn = div_rn(r.fy, (float)(p.H - 1));
  const float gxv = sub_rn(mul_rn(un, 2.f), 1.f), gyv = sub_rn(mul_rn(vn, 2.f), 1.f);
  const int S = p.S;
#pragma unroll 1
  for (int s = 0; s < S; ++s) {
    const SamplePoint q = sample_point(p, r, s);
    float vox[8];
    float f[V][CF + 7];
    const int cnt = gather_sample_regs<CF, V>(p, cams, s_view, s_tar_c, q.x, q.y, q.zz, gxv, gyv, q.dn, vox, f);
    const float4 o = nerf_mlp_eval<F, V>(sw, vox, f);
    const int64_t si = li * S + s;
    reinterpret_cast<float4*>(rp.raw)[si] = o;
    if (p.z_vals) p.z_vals[si] = q.z;
    if (p.vis_mask) p.vis_mask[si] = div_rn((float)cnt, (float)V);
    if (p.vis_count) p.vis_count[si] = cnt;
  }
}

template <int CF, int V>
static int launch_render(const bmv_render_rays_params& rp, cudaStream_t st) {
  using L = MlpLayout<CF + 3>;
  const size_t smem = (size_t)L::TOTAL * sizeof(float);
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(render_rays_kernel<CF, V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("bmv_render_rays: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  render_rays_kernel<CF, V><<<(unsigned)ceil_div64(rp.g.n_rays, kRenderThreads), kRenderThreads, smem, st>>>(rp);
  return check_launch("bmv_render_rays");
}

}  // namespace bmv

extern "C" BMV_API int bmv_render_rays(const bmv_render_rays_params* rp, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_render_rays");
  using namespace bmv;
  BMV_REQUIRE(rp != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays: null params");
  const bmv_raygen_fetch_params* p = &rp->g;
  BMV_REQUIRE(p->n_rays >= 0 && p->ray_begin >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays: bad ray range");
  if (p->n_rays == 0) return BMV_OK;
  BMV_REQUIRE(!p->xyz_in, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays: pointwise mode is not supported here");
  BMV_REQUIRE(p->rays12_in || (p->depth && p->std && p->near_far && (p->rays || p->ray_gen)), BMV_ERR_INVALID_ARGUMENT,
              "bmv_render_rays: null ray inputs");
  BMV_REQUIRE(p->volume && p->im_feat && p->rgb && p->src_exts && p->src_ixts && p->src_centers && p->tar_center &&
                  rp->mlp_weights && rp->raw,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays: null device pointer");
  BMV_REQUIRE(((uintptr_t)rp->mlp_weights & 15) == 0 && ((uintptr_t)rp->raw & 15) == 0, BMV_ERR_INVALID_ARGUMENT,
              "bmv_render_rays: weights/raw must be 16-byte aligned");
  BMV_REQUIRE(p->S >= 1 && (p->S == 1 || p->t), BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays: bad S / t");
  BMV_REQUIRE(p->H >= 2 && p->W >= 2 && p->hv >= 1 && p->wv >= 1 && p->Hf >= 2 && p->Wf >= 2 && p->Dv >= 1,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays: bad grid size");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->Cv == 8 && p->Cf == 8 && p->V == 3) return launch_render<8, 3>(*rp, st);
  set_error("bmv_render_rays: (Cv=%d, Cf=%d, V=%d) not instantiated (available: Cv=8, Cf=8, V=3)", p->Cv, p->Cf, p->V);
  return BMV_ERR_UNSUPPORTED_SHAPE;
}

extern "C" BMV_API int bmv_render_rays_supported(int Cv, int Cf, int V) { return Cv == 8 && Cf == 8 && V == 3; }
