// K3+K5 for ALL K cost-volume chains of a frame in ONE persistent launch, with a lean gather.
//
// Same result contract as bmv_render_rays_mma per chain (bit-exact z / visibility, split-fp16 tensor-core MLP), but:
//   * one launch walks the (chain, 32-sample round) work units of every chain (reference
//     lib/networks/boost_enerf/network.py:212-222 renders the chains one after the other): no per-chain launch tails,
//     the packed MLP weights are staged once per CTA instead of once per chain and CTA;
//   * the view ids of every chain are read from DEVICE memory, so a captured CUDA graph stays valid when the selected
//     triples change from frame to frame (sequence rendering, BASELINE config 5);
//   * the gather is rewritten around what ncu / SASS showed in render_mma.cu (2.9 k of 6.8 k warp instructions per
//     32-sample round): 32-bit indices and offsets (the 64-bit `si / S`, `ri % W` divisions alone were ~280
//     instructions), branch-free trilinear taps (validity folded into the weights), one tap set shared by the feature
//     and colour fetch of a view, exact `/ 2` written as `* 0.5`, reciprocals of loop-invariant divisors hoisted, and a
//     conservative filter in front of the bit-exact visibility test: the IEEE-division sequence of the reference
//     (lib/networks/enerf/utils.py:490-520) is only evaluated for the ~1e-5 of samples whose approximate NDC
//     coordinates lie within 1e-5 of a frustum edge — for all others the decision is provably the same.
// Layout requirements (checked on the host, otherwise BMV_ERR_UNSUPPORTED_SHAPE): Cv = Cf = 8, V = 3, dense
// channels-last volume / feature maps, (N,H,W,4) colours, everything < 2^31 elements.
#include <stdlib.h>

#include "lean_gather.cuh"
#include "mlp_mma_tile.cuh"
#include "render_multi_check.cuh"

namespace bmv {

constexpr int kRmWarps = 16;

// dbg (environment BMV_RM_DEBUG, measurement only: WRONG results): bit 0 skips the gather, bit 1 skips the MLP tiles
template <bool GEN, bool INV>
__global__ void __launch_bounds__(kRmWarps * 32, 1) render_multi_kernel(bmv_render_multi_params mp, int dbg) {
  constexpr int V = 3;
  const bmv_raygen_fetch_params& p = mp.g;
  extern __shared__ __align__(16) uint32_t smem_u[];
  uint32_t* sW = smem_u;
  const float* sV = reinterpret_cast<const float*>(smem_u);
  float* stage_all = reinterpret_cast<float*>(smem_u + MMA_PACK_WORDS);
  __shared__ LeanCam cams[BMV_MAX_VIEWS];
  __shared__ float s_tar_c[4];
  __shared__ int s_view[BMV_MAX_VOLUMES * V];
  for (int i = threadIdx.x * 4; i < MMA_PACK_WORDS; i += blockDim.x * 4)
    *reinterpret_cast<uint4*>(sW + i) = __ldg(reinterpret_cast<const uint4*>(mp.mlp_weights) + i / 4);
  if (threadIdx.x < mp.K * V) {                            // device-resident ids are clamped: they index shared memory
    const int id = mp.views ? __ldg(mp.views + threadIdx.x) : mp.views_host[threadIdx.x];
    s_view[threadIdx.x] = min(max(id, 0), mp.n_views - 1);
  }
  if (threadIdx.x < mp.n_views) cams[threadIdx.x] = lean_cam_load(p, threadIdx.x);
  if (threadIdx.x < 3) s_tar_c[threadIdx.x] = p.tar_center[threadIdx.x];
  __syncthreads();

  const RmCtx c = lean_ctx(p, mp.nf_plane_stride);

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* stage = stage_all + warp * 32 * kStageStride;
  const uint32_t S = (uint32_t)p.S;
  const uint32_t n_samples = (uint32_t)(p.n_rays * p.S);
  const uint32_t rounds = (n_samples + 31u) / 32u;          // per chain
  const uint32_t total = rounds * (uint32_t)mp.K;
  const float ba = sV[V_SC], bs = sV[V_SC + 1], b2 = sV[V_SC + 2];
  const bool unit_scale = p.render_scale == 1.f;
  const int vol_row0 = mp.vol_row0, map_row0 = mp.map_row0;

  // Warps free-run over the work units (gathers are LSU-bound, the MLP is tensor / issue bound: they overlap).
  for (uint32_t u = blockIdx.x * kRmWarps + warp; u < total; u += gridDim.x * kRmWarps) {
    const uint32_t k = u / rounds;                          // chain (warp-uniform)
    const uint32_t base = (u - k * rounds) * 32u;           // first sample of the round within the chain
    const int* views = s_view + k * V;
    const float* depth_k = mp.depth + (int64_t)k * mp.depth_k_stride;
    const float* std_k = mp.std + (int64_t)k * mp.std_k_stride;
    const float* nf_k = mp.near_far + (int64_t)k * mp.nf_k_stride;
    const float* vol_k = mp.volume + (int64_t)k * mp.vol_k_stride;
    const int64_t out0 = (int64_t)k * n_samples;
    // ------------------------------------------------------------ gather: one lane per sample
    if (!(dbg & 1)) {
      const uint32_t si_raw = base + lane;
      const bool live = si_raw < n_samples;
      const uint32_t si = live ? si_raw : n_samples - 1u;
      const uint32_t li = S == 2u ? (si >> 1) : si / S;
      const int s = (int)(si - li * S);
      const LeanPoint q = lean_sample_point<GEN, INV>(p, c, (uint32_t)p.ray_begin + li, s, depth_k, std_k, nf_k, map_row0);
      const float z = q.z;
      float* row = stage + lane * kStageStride;
      {
        float vox[8];
        lean_vox_fetch(p, c, vol_k, vol_row0, q, vox);
        *reinterpret_cast<float4*>(row) = make_float4(vox[0], vox[1], vox[2], vox[3]);
        *reinterpret_cast<float4*>(row + 4) = make_float4(vox[4], vox[5], vox[6], vox[7]);
      }
      // ---- per view: visibility, projection, bilinear feature + colour fetch (border padding), direction features
      int cnt = 0;
      const float3 tt = lean_target_dir(q, s_tar_c);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int view = views[v];
        const LeanCam& cam = cams[view];
        const LeanTaps tp = lean_project(p, c, cam, q, unit_scale);
        cnt += tp.visible ? 1 : 0;
        float f[16];
        lean_fetch_feat(p, view, tp, f);
        lean_fetch_rgb(p, view, tp, f + 8);
        lean_dir_feat(cam, q, tt, f + 11);
        f[15] = 0.f;
        float* fo = row + 8 + v * 16;
        *reinterpret_cast<float4*>(fo) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(fo + 4) = make_float4(f[4], f[5], f[6], f[7]);
        *reinterpret_cast<float4*>(fo + 8) = make_float4(f[8], f[9], f[10], f[11]);
        *reinterpret_cast<float4*>(fo + 12) = make_float4(f[12], f[13], f[14], f[15]);
      }
      if (live) {
        const int64_t oi = out0 + si;
        if (mp.z_vals) mp.z_vals[oi] = z;
        if (mp.vis_mask) mp.vis_mask[oi] = lean_vis_score3(cnt);
        if (mp.vis_count) mp.vis_count[oi] = cnt;
      }
    }
    __syncwarp();
    // ------------------------------------------------------------ MLP on two 16-sample tiles
#pragma unroll 1
    for (int tile = 0; tile < 2 && !(dbg & 2); ++tile) {
      const float* row0 = stage + (tile * 16 + g) * kStageStride;
      const float* row1 = row0 + 8 * kStageStride;
      float4 o[2];
      mlp_mma_tile(row0, row1, sW, sV, lane, ba, bs, b2, o);
      if (t == 0) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const uint32_t si = base + tile * 16 + g + 8 * r;
          if (si < n_samples) reinterpret_cast<float4*>(mp.raw)[out0 + si] = o[r];
        }
      }
    }
    __syncwarp();                                           // staging rows are rewritten next round
  }
}

}  // namespace bmv

extern "C" BMV_API int bmv_render_rays_multi(const bmv_render_multi_params* mp, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_render_rays_multi");
  using namespace bmv;
  if (const int rc = render_multi_validate(mp, "bmv_render_rays_multi"); rc != BMV_OK) return rc;
  const bmv_raygen_fetch_params* p = &mp->g;
  if (p->n_rays == 0) return BMV_OK;
  const size_t smem = (size_t)(MMA_PACK_WORDS + kRmWarps * 32 * kStageStride) * 4;
  const bool gen = p->ray_gen && !p->rays, invd = p->depth_inv != 0;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(render_multi_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(render_multi_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(render_multi_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(render_multi_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("bmv_render_rays_multi: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const int64_t units = ceil_div64(p->n_rays * p->S, 32) * mp->K;
  const int64_t want = ceil_div64(units, kRmWarps);
  const unsigned blocks = (unsigned)(want < kNumSMs ? want : kNumSMs);   // persistent: one CTA per SM
  cudaStream_t st = (cudaStream_t)stream;
  static const int dbg = getenv("BMV_RM_DEBUG") ? atoi(getenv("BMV_RM_DEBUG")) : 0;
  if (gen && invd) render_multi_kernel<true, true><<<blocks, kRmWarps * 32, smem, st>>>(*mp, dbg);
  else if (gen) render_multi_kernel<true, false><<<blocks, kRmWarps * 32, smem, st>>>(*mp, dbg);
  else if (invd) render_multi_kernel<false, true><<<blocks, kRmWarps * 32, smem, st>>>(*mp, dbg);
  else render_multi_kernel<false, false><<<blocks, kRmWarps * 32, smem, st>>>(*mp, dbg);
  return check_launch("bmv_render_rays_multi");
}
