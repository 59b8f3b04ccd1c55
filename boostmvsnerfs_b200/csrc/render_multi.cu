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
#include "mlp_mma_tile.cuh"
#include "raygen_common.cuh"

namespace bmv {

constexpr int kRmWarps = 16;

struct __align__(16) LeanCam {
  float4 E0, E1, E2;   // world->cam rows (r0 r1 r2 | t)
  float4 K0, K1, K2;   // intrinsics rows (w unused)
  float4 S0, S1;       // intrinsics rows 0, 1 multiplied by render_scale (k[:, :2] *= scale, one rounding each)
  float4 c;            // camera centre
};

// loop-invariant scalars of a launch (registers / uniform registers)
struct RmCtx {
  float isx, isy;            // (W-1), (H-1): visibility normalisation (inv_scale of the render grid)
  float r_isx, r_isy;        // their reciprocals (approximate filter only)
  float r_wf, r_hf;          // 1/(Wf-1), 1/(Hf-1)
  float up_sy, up_sx;        // align_corners upsample scales of the depth maps
  float wf1, hf1, wv1, hv1, dv1;
  int hwv;
};

__device__ __forceinline__ float frcp(float x) { return __frcp_rn(x); }
__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float4 lds4(const float4* q) { return *q; }

// The reference's own inside test on the pixel coordinates: two successive IEEE divisions per coordinate
// (lib/networks/enerf/utils.py:503-504, 514-516).  Out of line: executed for ~1e-5 of the samples.
__device__ __noinline__ bool visible_exact(float qx, float qy, float qz, float isx, float isy) {
  const float u = div_rn(div_rn(qx, qz), isx);
  const float v = div_rn(div_rn(qy, qz), isy);
  return (u >= 0.f) && (u <= 1.f) && (v >= 0.f) && (v <= 1.f);
}

// Visibility of one point in one view given q = K (R x + t) computed with the reference's op sequence (so q is
// bit-identical to the reference's): exact decision of point_visible(), IEEE divisions only near an edge.
__device__ __forceinline__ bool lean_visible(const RmCtx& c, float qx, float qy, float qz, float rq) {
  if (!(qz > 0.f)) return false;                          // also NaN: the reference's `z > 0` is false
  if (qz < 1e-6f) return visible_exact(qx, qy, qz, c.isx, c.isy);   // rq is 1 / max(qz, 1e-6): not 1 / qz here
  // approximate u, v: |ua - u_ref| <= ~1e-6 |u_ref| (rcp.approx 1 ulp, two more roundings; u_ref itself carries two)
  const float ua = qx * rq * c.r_isx, va = qy * rq * c.r_isy;
  constexpr float EPS = 1e-5f;
  const bool sure_in = ua > EPS && ua < 1.f - EPS && va > EPS && va < 1.f - EPS;
  if (sure_in) return true;
  const bool sure_out = ua < -EPS || ua > 1.f + EPS || va < -EPS || va > 1.f + EPS;   // +-inf compare like huge values
  if (sure_out) return false;
  return visible_exact(qx, qy, qz, c.isx, c.isy);        // within EPS of an edge, or NaN
}

template <bool GEN, bool INV>
__global__ void __launch_bounds__(kRmWarps * 32, 1) render_multi_kernel(bmv_render_multi_params mp) {
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
  if (threadIdx.x < mp.n_views) {
    const int v = threadIdx.x;
    const float* E = p.src_exts + v * 16;
    const float* Kx = p.src_ixts + v * 9;
    LeanCam cm;
    cm.E0 = make_float4(E[0], E[1], E[2], E[3]);
    cm.E1 = make_float4(E[4], E[5], E[6], E[7]);
    cm.E2 = make_float4(E[8], E[9], E[10], E[11]);
    cm.K0 = make_float4(Kx[0], Kx[1], Kx[2], 0.f);
    cm.K1 = make_float4(Kx[3], Kx[4], Kx[5], 0.f);
    cm.K2 = make_float4(Kx[6], Kx[7], Kx[8], 0.f);
    const float rs = p.render_scale;
    cm.S0 = make_float4(mul_rn(Kx[0], rs), mul_rn(Kx[1], rs), mul_rn(Kx[2], rs), 0.f);
    cm.S1 = make_float4(mul_rn(Kx[3], rs), mul_rn(Kx[4], rs), mul_rn(Kx[5], rs), 0.f);
    cm.c = make_float4(p.src_centers[v * 3], p.src_centers[v * 3 + 1], p.src_centers[v * 3 + 2], 0.f);
    cams[v] = cm;
  }
  if (threadIdx.x < 3) s_tar_c[threadIdx.x] = p.tar_center[threadIdx.x];
  __syncthreads();

  RmCtx c;
  c.isx = (float)(p.W - 1); c.isy = (float)(p.H - 1);
  c.r_isx = frcp(c.isx); c.r_isy = frcp(c.isy);
  c.wf1 = (float)(p.Wf - 1); c.hf1 = (float)(p.Hf - 1);
  c.r_wf = frcp(c.wf1); c.r_hf = frcp(c.hf1);
  c.wv1 = (float)(p.wv - 1); c.hv1 = (float)(p.hv - 1); c.dv1 = (float)(p.Dv - 1);
  c.up_sy = up_scale(p.hv, p.H); c.up_sx = up_scale(p.wv, p.W);
  c.hwv = mp.nf_plane_stride ? (int)mp.nf_plane_stride : p.hv * p.wv;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  float* stage = stage_all + warp * 32 * kStageStride;
  const uint32_t S = (uint32_t)p.S;
  const uint32_t n_samples = (uint32_t)(p.n_rays * p.S);
  const uint32_t rounds = (n_samples + 31u) / 32u;          // per chain
  const uint32_t total = rounds * (uint32_t)mp.K;
  const float ba = sV[V_SC], bs = sV[V_SC + 1], b2 = sV[V_SC + 2];
  const int Wi = p.W, Hi = p.H, wv = p.wv, hv = p.hv, Dv = p.Dv, Wf = p.Wf, Hf = p.Hf;
  const int vsx = (int)p.vol_x_stride, vsy = (int)p.vol_y_stride, vsd = (int)p.vol_d_stride;
  const float rgb_sc = p.rgb_scale, rgb_sf = p.rgb_shift;
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
    {
      const uint32_t si_raw = base + lane;
      const bool live = si_raw < n_samples;
      const uint32_t si = live ? si_raw : n_samples - 1u;
      const uint32_t li = S == 2u ? (si >> 1) : si / S;
      const int s = (int)(si - li * S);
      // ---- ray (build_rays): origin, direction, pixel
      float ox, oy, oz, dx, dy, dz;
      int px, py;
      float fx, fy;
      const uint32_t ri = (uint32_t)p.ray_begin + li;
      if (GEN) {
        const double* G = p.ray_gen;
        const int gx = (int)(ri % (uint32_t)Wi), gy = (int)(ri / (uint32_t)Wi);
        const double dxp = (double)gx, dyp = (double)gy;
        dx = (float)(__fma_rn(dyp, __ldg(G + 6), __dmul_rn(dxp, __ldg(G + 3))) + __ldg(G + 9));
        dy = (float)(__fma_rn(dyp, __ldg(G + 7), __dmul_rn(dxp, __ldg(G + 4))) + __ldg(G + 10));
        dz = (float)(__fma_rn(dyp, __ldg(G + 8), __dmul_rn(dxp, __ldg(G + 5))) + __ldg(G + 11));
        ox = (float)__ldg(G); oy = (float)__ldg(G + 1); oz = (float)__ldg(G + 2);
        fx = (float)gx; fy = (float)gy;
      } else {
        const float4 ra = __ldg(reinterpret_cast<const float4*>(p.rays) + 2 * (int64_t)ri);
        const float4 rb = __ldg(reinterpret_cast<const float4*>(p.rays) + 2 * (int64_t)ri + 1);
        ox = ra.x; oy = ra.y; oz = ra.z; dx = ra.w; dy = rb.x; dz = rb.y; fx = rb.z; fy = rb.w;
      }
      px = min(max((int)fx, 0), Wi - 1);                   // .long(): truncation toward zero
      py = min(max((int)fy, 0), Hi - 1);
      // ---- depth interval of the pixel: align_corners upsample of depth / std / near_far, then the clamp
      float rn, rf, nf0, nf1;
      {
        const UpCoord uy = up_coord_scaled(py, hv, c.up_sy), ux = up_coord_scaled(px, wv, c.up_sx);
        const int r0 = (uy.i0 - map_row0) * wv, r1 = (uy.i1 - map_row0) * wv;    // rows relative to the slab
        const int o00 = r0 + ux.i0, o01 = r0 + ux.i1, o10 = r1 + ux.i0, o11 = r1 + ux.i1;
        auto up = [&](const float* m) {
          const float a = __ldg(m + o00), b = __ldg(m + o01), cc = __ldg(m + o10), d = __ldg(m + o11);
          const float top = add_rn(mul_rn(ux.l0, a), mul_rn(ux.l1, b));
          const float bot = add_rn(mul_rn(ux.l0, cc), mul_rn(ux.l1, d));
          return add_rn(mul_rn(uy.l0, top), mul_rn(uy.l1, bot));
        };
        const float dep = up(depth_k), sd = up(std_k);
        nf0 = up(nf_k);
        nf1 = up(nf_k + c.hwv);
        if (INV) {
          rn = add_rn(dep, sd); rf = sub_rn(dep, sd);
          rn = rn > nf0 ? nf0 : rn;
          rf = rf < nf1 ? nf1 : rf;
        } else {
          rn = sub_rn(dep, sd); rf = add_rn(dep, sd);
          rn = rn < nf0 ? nf0 : rn;
          rf = rf > nf1 ? nf1 : rf;
        }
      }
      // ---- sample_along_depth
      const float tt = (S == 1u) ? 0.5f : __ldg(p.t + s);
      const float z = add_rn(rn, mul_rn(sub_rn(rf, rn), tt));
      float x, y, zz, dn;
      if (INV) {
        const float iz = div_rn(1.f, fmaxf(z, 1e-6f));
        x = add_rn(ox, mul_rn(dx, iz)); y = add_rn(oy, mul_rn(dy, iz)); zz = add_rn(oz, mul_rn(dz, iz));
        dn = div_rn(sub_rn(nf0, z), fmaxf(sub_rn(nf0, nf1), 1e-6f));
      } else {
        x = add_rn(ox, mul_rn(dx, z)); y = add_rn(oy, mul_rn(dy, z)); zz = add_rn(oz, mul_rn(dz, z));
        dn = div_rn(sub_rn(z, nf0), fmaxf(sub_rn(nf1, nf0), 1e-6f));
      }
      float* row = stage + lane * kStageStride;
      // ---- trilinear fetch of the regularised volume (zeros padding): validity folded into the weights
      {
        const float un = div_rn(fx, c.isx), vn = div_rn(fy, c.isy);
        const float gxv = sub_rn(mul_rn(un, 2.f), 1.f), gyv = sub_rn(mul_rn(vn, 2.f), 1.f);
        const float gz = sub_rn(mul_rn(dn, 2.f), 1.f);
        const float ix = unnormalize_ac(gxv, wv), iy = unnormalize_ac(gyv, hv), iz = unnormalize_ac(gz, Dv);
        const bool fin = coord_ok(ix) && coord_ok(iy) && coord_ok(iz);
        const float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
        const float fx1 = ix - x0, fy1 = iy - y0, fz1 = iz - z0;
        const float fx0 = (x0 + 1.f) - ix, fy0 = (y0 + 1.f) - iy, fz0 = (z0 + 1.f) - iz;
        // per-axis validity of the low / high corner (the reference skips out-of-range corners)
        const bool vx0 = fin && x0 >= 0.f && x0 <= c.wv1, vx1 = fin && x0 + 1.f >= 0.f && x0 + 1.f <= c.wv1;
        const bool vy0 = fin && y0 >= 0.f && y0 <= c.hv1, vy1 = fin && y0 + 1.f >= 0.f && y0 + 1.f <= c.hv1;
        const bool vz0 = fin && z0 >= 0.f && z0 <= c.dv1, vz1 = fin && z0 + 1.f >= 0.f && z0 + 1.f <= c.dv1;
        const float wx[2] = {vx0 ? fx0 : 0.f, vx1 ? fx1 : 0.f};
        const float wy[2] = {vy0 ? fy0 : 0.f, vy1 ? fy1 : 0.f};
        const float wz[2] = {vz0 ? fz0 : 0.f, vz1 ? fz1 : 0.f};
        // clamped integer corners (an invalid corner reads a valid address with weight 0)
        const float xc = fin ? fminf(fmaxf(x0, 0.f), c.wv1) : 0.f, yc = fin ? fminf(fmaxf(y0, 0.f), c.hv1) : 0.f;
        const float zc = fin ? fminf(fmaxf(z0, 0.f), c.dv1) : 0.f;
        const int xi = (int)xc, yi = (int)yc, zi = (int)zc;
        const int ox1 = (vx1 && vx0) ? vsx : 0, oy1 = (vy1 && vy0) ? vsy : 0, oz1 = (vz1 && vz0) ? vsd : 0;
        // when only the HIGH corner of an axis is valid (x0 = -1) the clamped index already is that corner
        const float* b000 = vol_k + (zi * vsd + (yi - vol_row0) * vsy + xi * vsx);
        float vox[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) vox[q] = 0.f;
#pragma unroll
        for (int corner = 0; corner < 8; ++corner) {
          const int bx = corner & 1, by = (corner >> 1) & 1, bz = corner >> 2;
          // the reference multiplies (wx * wy) * wz in this order
          const float w = (wx[bx] * wy[by]) * wz[bz];
          const float* src = b000 + ((bx ? ox1 : 0) + (by ? oy1 : 0) + (bz ? oz1 : 0));
          const float4 a = ldg4(src), b = ldg4(src + 4);
          vox[0] = fmaf(w, a.x, vox[0]); vox[1] = fmaf(w, a.y, vox[1]); vox[2] = fmaf(w, a.z, vox[2]); vox[3] = fmaf(w, a.w, vox[3]);
          vox[4] = fmaf(w, b.x, vox[4]); vox[5] = fmaf(w, b.y, vox[5]); vox[6] = fmaf(w, b.z, vox[6]); vox[7] = fmaf(w, b.w, vox[7]);
        }
        *reinterpret_cast<float4*>(row) = make_float4(vox[0], vox[1], vox[2], vox[3]);
        *reinterpret_cast<float4*>(row + 4) = make_float4(vox[4], vox[5], vox[6], vox[7]);
      }
      // ---- per view: visibility, projection, bilinear feature + colour fetch (border padding), direction features
      int cnt = 0;
      float ttx = sub_rn(x, s_tar_c[0]), tty = sub_rn(y, s_tar_c[1]), ttz = sub_rn(zz, s_tar_c[2]);
      {
        const float n = sqrt_approx(ttx * ttx + tty * tty + ttz * ttz) + 1e-6f;
        const float rinv = rcp_approx(n);
        ttx *= rinv; tty *= rinv; ttz *= rinv;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const int view = views[v];
        const LeanCam& cam = cams[view];
        // camera coordinates: the visibility path (bmm, then += t) and the fetch path (homogeneous matmul) round
        // identically — fma(1, t, acc) == acc + t — so one evaluation serves both; likewise q.z (unscaled K row 2)
        const float4 e0 = cam.E0, e1 = cam.E1, e2 = cam.E2;
        const float cx = dot4_gemm(x, y, zz, 1.f, e0.x, e0.y, e0.z, e0.w);
        const float cy = dot4_gemm(x, y, zz, 1.f, e1.x, e1.y, e1.z, e1.w);
        const float cz = dot4_gemm(x, y, zz, 1.f, e2.x, e2.y, e2.z, e2.w);
        const float4 k0 = cam.K0, k1 = cam.K1, k2 = cam.K2;
        const float vqx = dot3_gemm(cx, cy, cz, k0.x, k0.y, k0.z);
        const float vqy = dot3_gemm(cx, cy, cz, k1.x, k1.y, k1.z);
        const float qz = dot3_gemm(cx, cy, cz, k2.x, k2.y, k2.z);
        const float qzc = (qz != qz) ? qz : fmaxf(qz, 1e-6f);
        const float rq = rcp_approx(qzc);
        cnt += lean_visible(c, vqx, vqy, qz, rq) ? 1 : 0;
        float qx = vqx, qy = vqy;                           // render_scale == 1: the scaled intrinsics ARE the intrinsics
        if (!unit_scale) {
          const float4 s0 = cam.S0, s1 = cam.S1;
          qx = dot3_gemm(cx, cy, cz, s0.x, s0.y, s0.z);
          qy = dot3_gemm(cx, cy, cz, s1.x, s1.y, s1.z);
        }
        // grid = (pix / (W-1, H-1)) * 2 - 1, then ATen's ((g + 1) / 2) * (size - 1); divisions by reciprocal (<= 2 ulp)
        float gx = sub_rn(mul_rn(qx * rq * c.r_wf, 2.f), 1.f);
        float gy = sub_rn(mul_rn(qy * rq * c.r_hf, 2.f), 1.f);
        float ixf = unnormalize_ac(gx, Wf), iyf = unnormalize_ac(gy, Hf);
        ixf = fminf(c.wf1, fmaxf(ixf, 0.f));
        iyf = fminf(c.hf1, fmaxf(iyf, 0.f));
        const float x0 = floorf(ixf), y0 = floorf(iyf);
        const float wx1 = ixf - x0, wx0 = (x0 + 1.f) - ixf, wy1 = iyf - y0, wy0 = (y0 + 1.f) - iyf;
        const bool vx1 = x0 + 1.f <= c.wf1, vy1 = y0 + 1.f <= c.hf1;
        const int ix0 = (int)x0, iy0 = (int)y0;
        const int p00 = iy0 * Wf + ix0, dxp = vx1 ? 1 : 0, dyp = vy1 ? Wf : 0;
        const float w00 = wx0 * wy0, w01 = vx1 ? wx1 * wy0 : 0.f, w10 = vy1 ? wx0 * wy1 : 0.f, w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
        const float* fm = p.im_feat + (int64_t)view * p.imf_view_stride + p00 * 8;
        const float* fr = p.rgb + (int64_t)view * p.rgb_view_stride + p00 * 4;
        float f[16];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float4 a = ldg4(fm + 4 * h), b = ldg4(fm + dxp * 8 + 4 * h);
          const float4 cc = ldg4(fm + dyp * 8 + 4 * h), d = ldg4(fm + (dyp + dxp) * 8 + 4 * h);
          f[4 * h + 0] = fmaf(w11, d.x, fmaf(w10, cc.x, fmaf(w01, b.x, w00 * a.x)));
          f[4 * h + 1] = fmaf(w11, d.y, fmaf(w10, cc.y, fmaf(w01, b.y, w00 * a.y)));
          f[4 * h + 2] = fmaf(w11, d.z, fmaf(w10, cc.z, fmaf(w01, b.z, w00 * a.z)));
          f[4 * h + 3] = fmaf(w11, d.w, fmaf(w10, cc.w, fmaf(w01, b.w, w00 * a.w)));
        }
        {
          const float4 a = ldg4(fr), b = ldg4(fr + dxp * 4), cc = ldg4(fr + dyp * 4), d = ldg4(fr + (dyp + dxp) * 4);
          f[8] = fmaf(w11, fmaf(d.x, rgb_sc, rgb_sf), fmaf(w10, fmaf(cc.x, rgb_sc, rgb_sf), fmaf(w01, fmaf(b.x, rgb_sc, rgb_sf), w00 * fmaf(a.x, rgb_sc, rgb_sf))));
          f[9] = fmaf(w11, fmaf(d.y, rgb_sc, rgb_sf), fmaf(w10, fmaf(cc.y, rgb_sc, rgb_sf), fmaf(w01, fmaf(b.y, rgb_sc, rgb_sf), w00 * fmaf(a.y, rgb_sc, rgb_sf))));
          f[10] = fmaf(w11, fmaf(d.z, rgb_sc, rgb_sf), fmaf(w10, fmaf(cc.z, rgb_sc, rgb_sf), fmaf(w01, fmaf(b.z, rgb_sc, rgb_sf), w00 * fmaf(a.z, rgb_sc, rgb_sf))));
        }
        {
          const float4 cc = cam.c;
          float sx = sub_rn(x, cc.x), sy = sub_rn(y, cc.y), sz = sub_rn(zz, cc.z);
          const float n = sqrt_approx(sx * sx + sy * sy + sz * sz) + 1e-6f;
          const float rinv = rcp_approx(n);
          sx *= rinv; sy *= rinv; sz *= rinv;
          const float ex = sub_rn(ttx, sx), ey = sub_rn(tty, sy), ez = sub_rn(ttz, sz);
          const float en = fmaxf(sqrt_approx(ex * ex + ey * ey + ez * ez), 1e-6f);
          const float re = rcp_approx(en);
          f[11] = ex * re; f[12] = ey * re; f[13] = ez * re;
          f[14] = ttx * sx + tty * sy + ttz * sz;
          f[15] = 0.f;
        }
        float* fo = row + 8 + v * 16;
        *reinterpret_cast<float4*>(fo) = make_float4(f[0], f[1], f[2], f[3]);
        *reinterpret_cast<float4*>(fo + 4) = make_float4(f[4], f[5], f[6], f[7]);
        *reinterpret_cast<float4*>(fo + 8) = make_float4(f[8], f[9], f[10], f[11]);
        *reinterpret_cast<float4*>(fo + 12) = make_float4(f[12], f[13], f[14], f[15]);
      }
      if (live) {
        const int64_t oi = out0 + si;
        if (mp.z_vals) mp.z_vals[oi] = z;
        // RN(cnt / 3) for cnt = 0..3 (the reference's `m /= V`): constants instead of an IEEE division
        if (mp.vis_mask) mp.vis_mask[oi] = cnt == 0 ? 0.f : (cnt == 1 ? 0.333333343267440796f : (cnt == 2 ? 0.666666686534881592f : 1.f));
        if (mp.vis_count) mp.vis_count[oi] = cnt;
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
  BMV_REQUIRE(mp != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: null params");
  const bmv_raygen_fetch_params* p = &mp->g;
  BMV_REQUIRE(p->n_rays >= 0 && p->ray_begin >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: bad ray range");
  BMV_REQUIRE(mp->K >= 1 && mp->K <= BMV_MAX_VOLUMES && mp->n_views >= 1 && mp->n_views <= BMV_MAX_VIEWS,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: K must be 1..%d and n_views 1..%d", BMV_MAX_VOLUMES, BMV_MAX_VIEWS);
  if (p->n_rays == 0) return BMV_OK;
  BMV_REQUIRE(mp->depth && mp->std && mp->near_far && mp->volume && (p->rays || p->ray_gen) && p->im_feat && p->rgb &&
                  p->src_exts && p->src_ixts && p->src_centers && p->tar_center && mp->mlp_weights && mp->raw,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: null device pointer");
  BMV_REQUIRE(((uintptr_t)mp->mlp_weights & 15) == 0 && ((uintptr_t)mp->raw & 15) == 0 && ((uintptr_t)mp->volume & 15) == 0 &&
                  ((uintptr_t)p->im_feat & 15) == 0 && ((uintptr_t)p->rgb & 15) == 0 && (!p->rays || ((uintptr_t)p->rays & 15) == 0),
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: pointers must be 16-byte aligned");
  BMV_REQUIRE(p->S >= 1 && (p->S == 1 || p->t), BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: bad S / t");
  BMV_REQUIRE(p->H >= 2 && p->W >= 2 && p->hv >= 1 && p->wv >= 1 && p->Hf >= 2 && p->Wf >= 2 && p->Dv >= 1,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: bad grid size");
  BMV_REQUIRE(p->Cv == 8 && p->Cf == 8 && p->V == 3, BMV_ERR_UNSUPPORTED_SHAPE,
              "bmv_render_rays_multi: (Cv=%d, Cf=%d, V=%d) not instantiated (8, 8, 3)", p->Cv, p->Cf, p->V);
  // dense channels-last layouts, 32-bit offsets
  BMV_REQUIRE(p->vol_c_stride == 1 && p->vol_x_stride == 8 && p->vol_y_stride == (int64_t)p->wv * 8 &&
                  p->vol_d_stride % 4 == 0 && p->vol_d_stride > 0 && p->vol_d_stride <= (int64_t)p->hv * p->wv * 8 &&
                  mp->vol_k_stride % 4 == 0,
              BMV_ERR_UNSUPPORTED_SHAPE, "bmv_render_rays_multi: the volumes must be (D,rows,w,8) channels-last with dense rows");
  BMV_REQUIRE(mp->vol_row0 >= 0 && mp->vol_row0 < p->hv && mp->map_row0 >= 0 && mp->map_row0 < p->hv && mp->nf_plane_stride >= 0,
              BMV_ERR_INVALID_ARGUMENT, "bmv_render_rays_multi: bad slab rows");
  BMV_REQUIRE(p->imf_c_stride == 1 && p->imf_x_stride == 8 && p->imf_y_stride == (int64_t)p->Wf * 8 && p->imf_view_stride % 4 == 0,
              BMV_ERR_UNSUPPORTED_SHAPE, "bmv_render_rays_multi: im_feat must be dense (N,Hf,Wf,8) channels-last");
  BMV_REQUIRE(p->rgb_c_stride == 1 && p->rgb_x_stride == 4 && p->rgb_y_stride == (int64_t)p->Wf * 4 && p->rgb_view_stride % 4 == 0,
              BMV_ERR_UNSUPPORTED_SHAPE, "bmv_render_rays_multi: rgb must be dense (N,Hf,Wf,4)");
  BMV_REQUIRE((int64_t)p->Dv * p->vol_d_stride < (1ll << 31) && (int64_t)p->Hf * p->Wf * 8 < (1ll << 31) &&
                  p->n_rays * p->S < (1ll << 31) && p->ray_begin + p->n_rays < (1ll << 31),
              BMV_ERR_UNSUPPORTED_SHAPE, "bmv_render_rays_multi: tensors too large for 32-bit offsets");
  for (int i = 0; !mp->views && i < mp->K * 3; ++i)
    BMV_REQUIRE(mp->views_host[i] >= 0 && mp->views_host[i] < mp->n_views, BMV_ERR_INVALID_ARGUMENT,
                "bmv_render_rays_multi: view id %d out of range", mp->views_host[i]);
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
  if (gen && invd) render_multi_kernel<true, true><<<blocks, kRmWarps * 32, smem, st>>>(*mp);
  else if (gen) render_multi_kernel<true, false><<<blocks, kRmWarps * 32, smem, st>>>(*mp);
  else if (invd) render_multi_kernel<false, true><<<blocks, kRmWarps * 32, smem, st>>>(*mp);
  else render_multi_kernel<false, false><<<blocks, kRmWarps * 32, smem, st>>>(*mp);
  return check_launch("bmv_render_rays_multi");
}
