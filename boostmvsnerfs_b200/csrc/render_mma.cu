// K3+K5 on tensor cores: the fused gather + per-sample MLP of render_fused.cu with every Linear layer
// evaluated by warp-level MMAs (mma.sync m16n8k16, fp16 operands, fp32 accumulate) instead of fp32
// FMAs.  ncu on the fp32-FMA kernel (profiles/round1c): 790 M warp instructions per launch, FMA pipe
// 49 %, LDS-latency bound at 12 warps/SM — 0.37 of the fp32 peak and 37 % of the frame.
//
// Accuracy: each fp32 operand x is split as x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
// (22 significant bits); a product uses three MMAs (hi*hi + hi*lo + lo*hi), accumulation is fp32.
// The dropped lo*lo term is 2^-22 relative; results agree with the fp32 path to ~1e-6, well inside
// the 1e-4 parity bar (no TF32/bf16-style precision loss).
//
// Data flow per warp and round (32 samples = two 16-row MMA tiles):
//   gather (one lane per sample, same arithmetic as render_fused.cu) -> staging rows in shared memory
//   -> per tile: Agg.view_fc (elementwise, fragment domain) -> global_fc (MMA) -> view softmax -> fc
//   (MMA) -> lr0 (MMA) -> sigma -> color.0 (MMA: shared part once, per-view part on top) -> color.2
//   -> view softmax -> rgb.  Layer outputs (C fragments) are re-used directly as the next layer's A
//   fragments: for m16n8k16 two adjacent 8-column C tiles form one 16-wide K tile with no data movement.
// Weights are pre-split and pre-arranged in B-fragment order by mlp_pack.pack_nerf_weights_mma(): one
// 16-byte shared-memory load per lane per (k-tile, n-tile) = {b0_hi, b1_hi, b0_lo, b1_lo}.
#include "mlp_mma_tile.cuh"
#include "raygen_common.cuh"

namespace bmv {

constexpr int kMmaWarps = 16;      // 128 registers per thread (76 B of spills) beat 12 warps at 166: 1.88 -> 1.74 ms per C2 frame
template <bool VEC>
__global__ void __launch_bounds__(kMmaWarps * 32, 1) render_rays_mma_kernel(bmv_render_rays_params rp) {
  constexpr int V = 3, CF = 8;
  const bmv_raygen_fetch_params& p = rp.g;
  extern __shared__ __align__(16) uint32_t smem_u[];
  uint32_t* sW = smem_u;
  const float* sV = reinterpret_cast<const float*>(smem_u);
  float* stage_all = reinterpret_cast<float*>(smem_u + MMA_PACK_WORDS);
  __shared__ ViewCam cams[V];
  __shared__ float s_tar_c[3];
  __shared__ int s_view[V];
  for (int i = threadIdx.x * 4; i < MMA_PACK_WORDS; i += blockDim.x * 4)
    *reinterpret_cast<uint4*>(sW + i) = __ldg(reinterpret_cast<const uint4*>(rp.mlp_weights) + i / 4);
  if (threadIdx.x < V) s_view[threadIdx.x] = p.view[threadIdx.x];
  __syncthreads();
  for (int v = 0; v < V; ++v) load_cam(&cams[v], p.src_exts, p.src_ixts, p.src_centers, s_view[v], threadIdx.x);
  if (threadIdx.x < 3) s_tar_c[threadIdx.x] = p.tar_center[threadIdx.x];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* stage = stage_all + warp * 32 * kStageStride;
  const int S = p.S;
  const int64_t n_samples = p.n_rays * S;
  const float ba = sV[V_SC], bs = sV[V_SC + 1], b2 = sV[V_SC + 2];

  // Warps run their rounds independently: while some gather (LSU-bound) others are in the MMA chain.  (With the
  // scalar-load gather the round body was ~90 KB of code and free-running warps thrashed the instruction cache;
  // with 16-byte loads the body is small enough that free-running is 3.5 % faster than a per-round barrier.)
  const int64_t per_round = (int64_t)gridDim.x * kMmaWarps * 32;
  const int64_t rounds = (n_samples + per_round - 1) / per_round;
  for (int64_t rd = 0; rd < rounds; ++rd) {
    const int64_t base = rd * per_round + ((int64_t)blockIdx.x * kMmaWarps + warp) * 32;
    if (base >= n_samples) continue;
    // ------------------------------------------------------------ gather: one lane per sample
    {
      const int64_t si = base + lane;
      float vox[8];
      float f[V][CF + 7];
      if (si < n_samples) {
        const int64_t li = si / S;
        const int s = (int)(si % S);
        const RaySetup r = ray_setup(p, li);
        const float un = div_rn(r.fx, (float)(p.W - 1)), vn = div_rn(r.fy, (float)(p.H - 1));
        const float gxv = sub_rn(mul_rn(un, 2.f), 1.f), gyv = sub_rn(mul_rn(vn, 2.f), 1.f);
        const SamplePoint q = sample_point(p, r, s);
        const int cnt = gather_sample_regs<CF, V, VEC>(p, cams, s_view, s_tar_c, q.x, q.y, q.zz, gxv, gyv, q.dn, vox, f);
        if (p.z_vals) p.z_vals[si] = q.z;
        if (p.vis_mask) p.vis_mask[si] = div_rn((float)cnt, (float)V);
        if (p.vis_count) p.vis_count[si] = cnt;
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) vox[c] = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v)
#pragma unroll
          for (int c = 0; c < CF + 7; ++c) f[v][c] = 0.f;
      }
      float* row = stage + lane * kStageStride;
      *reinterpret_cast<float4*>(row) = make_float4(vox[0], vox[1], vox[2], vox[3]);
      *reinterpret_cast<float4*>(row + 4) = make_float4(vox[4], vox[5], vox[6], vox[7]);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float* fr = row + 8 + v * 16;
        *reinterpret_cast<float4*>(fr) = make_float4(f[v][0], f[v][1], f[v][2], f[v][3]);
        *reinterpret_cast<float4*>(fr + 4) = make_float4(f[v][4], f[v][5], f[v][6], f[v][7]);
        *reinterpret_cast<float4*>(fr + 8) = make_float4(f[v][8], f[v][9], f[v][10], f[v][11]);
        *reinterpret_cast<float4*>(fr + 12) = make_float4(f[v][12], f[v][13], f[v][14], 0.f);
      }
    }
    __syncwarp();

    // ------------------------------------------------------------ MLP on two 16-sample tiles
#pragma unroll 1
    for (int tile = 0; tile < 2; ++tile) {
      const float* row0 = stage + (tile * 16 + g) * kStageStride;
      const float* row1 = row0 + 8 * kStageStride;
      float4 o[2];
      mlp_mma_tile(row0, row1, sW, sV, lane, ba, bs, b2, o);
      if (t == 0) {                                         // lane t == 0 of each quad writes rows g and g+8
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const int64_t si = base + tile * 16 + g + 8 * r;
          if (si < n_samples) reinterpret_cast<float4*>(rp.raw)[si] = o[r];
        }
      }
    }
    __syncwarp();                                           // staging rows are rewritten next round
  }
}

}  // namespace bmv

extern "C" BMV_API int bmv_render_rays_mma_weight_words(void) { return bmv::MMA_PACK_WORDS; }

extern "C" BMV_API int bmv_render_rays_mma(const bmv_render_rays_params* rp, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_render_rays_mma");
  using namespace bmv;
  BMV_REQUIRE(rp != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_mma: null params");
  const bmv_raygen_fetch_params* p = &rp->g;
  BMV_REQUIRE(p->n_rays >= 0 && p->ray_begin >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_mma: bad ray range");
  if (p->n_rays == 0) return BMV_OK;
  BMV_REQUIRE(!p->xyz_in, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_mma: pointwise mode is not supported here");
  BMV_REQUIRE(p->rays12_in || (p->depth && p->std && p->near_far && (p->rays || p->ray_gen)), BMV_ERR_INVALID_ARGUMENT,
              "bmv_render_rays_mma: null ray inputs");
  BMV_REQUIRE(p->volume && p->im_feat && p->rgb && p->src_exts && p->src_ixts && p->src_centers && p->tar_center &&
                  rp->mlp_weights && rp->raw,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_mma: null device pointer");
  BMV_REQUIRE(((uintptr_t)rp->mlp_weights & 15) == 0 && ((uintptr_t)rp->raw & 15) == 0, BMV_ERR_INVALID_ARGUMENT,
              "bmv_render_rays_mma: weights/raw must be 16-byte aligned");
  BMV_REQUIRE(p->S >= 1 && (p->S == 1 || p->t), BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_mma: bad S / t");
  BMV_REQUIRE(p->H >= 2 && p->W >= 2 && p->hv >= 1 && p->wv >= 1 && p->Hf >= 2 && p->Wf >= 2 && p->Dv >= 1,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_mma: bad grid size");
  BMV_REQUIRE(p->Cv == 8 && p->Cf == 8 && p->V == 3, BMV_ERR_UNSUPPORTED_SHAPE,
              "bmv_render_rays_mma: (Cv=%d, Cf=%d, V=%d) not instantiated (8, 8, 3)", p->Cv, p->Cf, p->V);
  const size_t smem = (size_t)(MMA_PACK_WORDS + kMmaWarps * 32 * kStageStride) * 4;
  static DeviceOnce configured;
  if (const int cfg_dev = configured.needed(); cfg_dev >= 0) {
    cudaError_t e = cudaFuncSetAttribute(render_rays_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(render_rays_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("bmv_render_rays_mma: cannot reserve %zu B shared memory: %s", smem, cudaGetErrorString(e));
      return BMV_ERR_CUDA_LAUNCH;
    }
    configured.done(cfg_dev);
  }
  const int64_t rounds = ceil_div64(p->n_rays * p->S, kMmaWarps * 32);
  const unsigned blocks = (unsigned)(rounds < kNumSMs ? rounds : kNumSMs);   // persistent: one CTA per SM
  if (gather_vec_ok(*p)) render_rays_mma_kernel<true><<<blocks, kMmaWarps * 32, smem, (cudaStream_t)stream>>>(*rp);
  else render_rays_mma_kernel<false><<<blocks, kMmaWarps * 32, smem, (cudaStream_t)stream>>>(*rp);
  return check_launch("bmv_render_rays_mma");
}
