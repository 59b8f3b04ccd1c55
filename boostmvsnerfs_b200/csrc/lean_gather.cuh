// The lean per-sample gather of the multi-chain render kernels (render_multi.cu: mma.sync MLP; render_multi_umma.cu:
// tcgen05 MLP) as inlined pieces: ray + depth interval + sample point, trilinear volume fetch, per-view projection /
// bilinear taps / visibility, direction features.  Arithmetic contract (reference lib/networks/enerf/utils.py:392-443,
// 458-460, 490-520, 753-786): z, the sample position and the visibility decision are bit-exact (separately rounded
// ops, IEEE divisions where a decision depends on them); fetch coordinates use reciprocals (<= 2 ulp).
// How the instruction count was cut (ncu / SASS of render_mma.cu: 2.9 k of 6.8 k warp instructions per 32-sample
// round): 32-bit indices and offsets, branch-free trilinear taps (validity folded into the weights), one tap set for
// the feature and colour fetch of a view, `/ 2` as `* 0.5`, hoisted reciprocals, and a conservative filter in front of
// the bit-exact visibility test (IEEE divisions only within 1e-5 of a frustum edge).
#pragma once
#include "raygen_common.cuh"

namespace bmv {

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

// one 32-byte channels-last texel (8 fp32 channels) with ONE 256-bit load (sm_100: LDG.E.ENL2.256): a scattered tap
// costs the L1 tag stage one request instead of two — the gather is bound by exactly that (ncu: l1tex 63 %, the
// largest consumer in both multi-chain render kernels).  32-byte aligned address.
__device__ __forceinline__ void ldg8(const float* q, float4& a, float4& b) {
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(q));
}

__device__ __forceinline__ LeanCam lean_cam_load(const bmv_raygen_fetch_params& p, int v) {
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
  return cm;
}

__device__ __forceinline__ RmCtx lean_ctx(const bmv_raygen_fetch_params& p, int64_t nf_plane_stride) {
  RmCtx c;
  c.isx = (float)(p.W - 1); c.isy = (float)(p.H - 1);
  c.r_isx = frcp(c.isx); c.r_isy = frcp(c.isy);
  c.wf1 = (float)(p.Wf - 1); c.hf1 = (float)(p.Hf - 1);
  c.r_wf = frcp(c.wf1); c.r_hf = frcp(c.hf1);
  c.wv1 = (float)(p.wv - 1); c.hv1 = (float)(p.hv - 1); c.dv1 = (float)(p.Dv - 1);
  c.up_sy = up_scale(p.hv, p.H); c.up_sx = up_scale(p.wv, p.W);
  c.hwv = nf_plane_stride ? (int)nf_plane_stride : p.hv * p.wv;
  return c;
}

// The reference's own inside test on the pixel coordinates: two successive IEEE divisions per coordinate
// (lib/networks/enerf/utils.py:503-504, 514-516).  Out of line: executed for ~1e-5 of the samples.
static __device__ __noinline__ bool visible_exact(float qx, float qy, float qz, float isx, float isy) {
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

// ---- ray (build_rays) + depth interval of the pixel + sample_along_depth: position of sample `s` of ray `ri`
struct LeanPoint { float x, y, zz, z, dn, fx, fy; };

template <bool GEN, bool INV>
__device__ __forceinline__ LeanPoint lean_sample_point(const bmv_raygen_fetch_params& p, const RmCtx& c, uint32_t ri, int s,
                                                       const float* __restrict__ depth_k, const float* __restrict__ std_k,
                                                       const float* __restrict__ nf_k, int map_row0) {
  const int Wi = p.W, Hi = p.H, wv = p.wv, hv = p.hv;
  float ox, oy, oz, dx, dy, dz;
  float fx, fy;
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
  const int px = min(max((int)fx, 0), Wi - 1);           // .long(): truncation toward zero
  const int py = min(max((int)fy, 0), Hi - 1);
  // depth interval of the pixel: align_corners upsample of depth / std / near_far, then the clamp
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
  // sample_along_depth
  const float tt = (p.S == 1) ? 0.5f : __ldg(p.t + s);
  LeanPoint q;
  q.fx = fx; q.fy = fy;
  q.z = add_rn(rn, mul_rn(sub_rn(rf, rn), tt));
  if (INV) {
    const float iz = div_rn(1.f, fmaxf(q.z, 1e-6f));
    q.x = add_rn(ox, mul_rn(dx, iz)); q.y = add_rn(oy, mul_rn(dy, iz)); q.zz = add_rn(oz, mul_rn(dz, iz));
    q.dn = div_rn(sub_rn(nf0, q.z), fmaxf(sub_rn(nf0, nf1), 1e-6f));
  } else {
    q.x = add_rn(ox, mul_rn(dx, q.z)); q.y = add_rn(oy, mul_rn(dy, q.z)); q.zz = add_rn(oz, mul_rn(dz, q.z));
    q.dn = div_rn(sub_rn(q.z, nf0), fmaxf(sub_rn(nf1, nf0), 1e-6f));
  }
  return q;
}

// ---- trilinear fetch of the regularised volume (zeros padding, dense channels-last (D,rows,w,8)): validity folded
// into the weights, an invalid corner reads a clamped (valid) address with weight 0
__device__ __forceinline__ void lean_vox_fetch(const bmv_raygen_fetch_params& p, const RmCtx& c, const float* __restrict__ vol_k,
                                               int vol_row0, const LeanPoint& q, float (&vox)[8]) {
  const int vsx = (int)p.vol_x_stride, vsy = (int)p.vol_y_stride, vsd = (int)p.vol_d_stride;
  const float un = div_rn(q.fx, c.isx), vn = div_rn(q.fy, c.isy);
  const float gxv = sub_rn(mul_rn(un, 2.f), 1.f), gyv = sub_rn(mul_rn(vn, 2.f), 1.f);
  const float gz = sub_rn(mul_rn(q.dn, 2.f), 1.f);
  const float ix = unnormalize_ac(gxv, p.wv), iy = unnormalize_ac(gyv, p.hv), iz = unnormalize_ac(gz, p.Dv);
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
  // clamped integer corners
  const float xc = fin ? fminf(fmaxf(x0, 0.f), c.wv1) : 0.f, yc = fin ? fminf(fmaxf(y0, 0.f), c.hv1) : 0.f;
  const float zc = fin ? fminf(fmaxf(z0, 0.f), c.dv1) : 0.f;
  const int xi = (int)xc, yi = (int)yc, zi = (int)zc;
  const int ox1 = (vx1 && vx0) ? vsx : 0, oy1 = (vy1 && vy0) ? vsy : 0, oz1 = (vz1 && vz0) ? vsd : 0;
  // when only the HIGH corner of an axis is valid (x0 = -1) the clamped index already is that corner
  const float* b000 = vol_k + (zi * vsd + (yi - vol_row0) * vsy + xi * vsx);
#pragma unroll
  for (int k = 0; k < 8; ++k) vox[k] = 0.f;
#pragma unroll
  for (int corner = 0; corner < 8; ++corner) {
    const int bx = corner & 1, by = (corner >> 1) & 1, bz = corner >> 2;
    const float w = (wx[bx] * wy[by]) * wz[bz];           // the reference multiplies (wx * wy) * wz in this order
    const float* src = b000 + ((bx ? ox1 : 0) + (by ? oy1 : 0) + (bz ? oz1 : 0));
    float4 a, b;
    ldg8(src, a, b);
    vox[0] = fmaf(w, a.x, vox[0]); vox[1] = fmaf(w, a.y, vox[1]); vox[2] = fmaf(w, a.z, vox[2]); vox[3] = fmaf(w, a.w, vox[3]);
    vox[4] = fmaf(w, b.x, vox[4]); vox[5] = fmaf(w, b.y, vox[5]); vox[6] = fmaf(w, b.z, vox[6]); vox[7] = fmaf(w, b.w, vox[7]);
  }
}

// ---- unit vector from the target camera centre to the sample
__device__ __forceinline__ float3 lean_target_dir(const LeanPoint& q, const float* tar_c) {
  float ttx = sub_rn(q.x, tar_c[0]), tty = sub_rn(q.y, tar_c[1]), ttz = sub_rn(q.zz, tar_c[2]);
  const float n = sqrt_approx(ttx * ttx + tty * tty + ttz * ttz) + 1e-6f;
  const float rinv = rcp_approx(n);
  return make_float3(ttx * rinv, tty * rinv, ttz * rinv);
}

// ---- one source view: projection (shared by the visibility test and the fetch), bilinear taps (border padding)
struct LeanTaps {
  int p00, dxp, dyp;            // pixel index of the top-left tap, +1 / +Wf where the neighbour exists
  float w00, w01, w10, w11;
  bool visible;
};

__device__ __forceinline__ LeanTaps lean_project(const bmv_raygen_fetch_params& p, const RmCtx& c, const LeanCam& cam,
                                                 const LeanPoint& q, bool unit_scale) {
  // camera coordinates: the visibility path (bmm, then += t) and the fetch path (homogeneous matmul) round
  // identically — fma(1, t, acc) == acc + t — so one evaluation serves both; likewise q.z (unscaled K row 2)
  const float4 e0 = cam.E0, e1 = cam.E1, e2 = cam.E2;
  const float cx = dot4_gemm(q.x, q.y, q.zz, 1.f, e0.x, e0.y, e0.z, e0.w);
  const float cy = dot4_gemm(q.x, q.y, q.zz, 1.f, e1.x, e1.y, e1.z, e1.w);
  const float cz = dot4_gemm(q.x, q.y, q.zz, 1.f, e2.x, e2.y, e2.z, e2.w);
  const float4 k0 = cam.K0, k1 = cam.K1, k2 = cam.K2;
  const float vqx = dot3_gemm(cx, cy, cz, k0.x, k0.y, k0.z);
  const float vqy = dot3_gemm(cx, cy, cz, k1.x, k1.y, k1.z);
  const float qz = dot3_gemm(cx, cy, cz, k2.x, k2.y, k2.z);
  const float qzc = (qz != qz) ? qz : fmaxf(qz, 1e-6f);
  const float rq = rcp_approx(qzc);
  LeanTaps t;
  t.visible = lean_visible(c, vqx, vqy, qz, rq);
  float qx = vqx, qy = vqy;                               // render_scale == 1: the scaled intrinsics ARE the intrinsics
  if (!unit_scale) {
    const float4 s0 = cam.S0, s1 = cam.S1;
    qx = dot3_gemm(cx, cy, cz, s0.x, s0.y, s0.z);
    qy = dot3_gemm(cx, cy, cz, s1.x, s1.y, s1.z);
  }
  // grid = (pix / (W-1, H-1)) * 2 - 1, then ATen's ((g + 1) / 2) * (size - 1); divisions by reciprocal (<= 2 ulp)
  const float gx = sub_rn(mul_rn(qx * rq * c.r_wf, 2.f), 1.f);
  const float gy = sub_rn(mul_rn(qy * rq * c.r_hf, 2.f), 1.f);
  float ixf = unnormalize_ac(gx, p.Wf), iyf = unnormalize_ac(gy, p.Hf);
  ixf = fminf(c.wf1, fmaxf(ixf, 0.f));
  iyf = fminf(c.hf1, fmaxf(iyf, 0.f));
  const float x0 = floorf(ixf), y0 = floorf(iyf);
  const float wx1 = ixf - x0, wx0 = (x0 + 1.f) - ixf, wy1 = iyf - y0, wy0 = (y0 + 1.f) - iyf;
  const bool vx1 = x0 + 1.f <= c.wf1, vy1 = y0 + 1.f <= c.hf1;
  const int ix0 = (int)x0, iy0 = (int)y0;
  t.p00 = iy0 * p.Wf + ix0; t.dxp = vx1 ? 1 : 0; t.dyp = vy1 ? p.Wf : 0;
  t.w00 = wx0 * wy0; t.w01 = vx1 ? wx1 * wy0 : 0.f; t.w10 = vy1 ? wx0 * wy1 : 0.f; t.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return t;
}

// 8 feature channels of a dense (N,Hf,Wf,8) map at the taps
__device__ __forceinline__ void lean_fetch_feat(const bmv_raygen_fetch_params& p, int view, const LeanTaps& t, float* f) {
  const float* fm = p.im_feat + (int64_t)view * p.imf_view_stride + t.p00 * 8;
  float4 a[2], b[2], cc[2], d[2];
  ldg8(fm, a[0], a[1]);
  ldg8(fm + t.dxp * 8, b[0], b[1]);
  ldg8(fm + t.dyp * 8, cc[0], cc[1]);
  ldg8(fm + (t.dyp + t.dxp) * 8, d[0], d[1]);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    f[4 * h + 0] = fmaf(t.w11, d[h].x, fmaf(t.w10, cc[h].x, fmaf(t.w01, b[h].x, t.w00 * a[h].x)));
    f[4 * h + 1] = fmaf(t.w11, d[h].y, fmaf(t.w10, cc[h].y, fmaf(t.w01, b[h].y, t.w00 * a[h].y)));
    f[4 * h + 2] = fmaf(t.w11, d[h].z, fmaf(t.w10, cc[h].z, fmaf(t.w01, b[h].z, t.w00 * a[h].z)));
    f[4 * h + 3] = fmaf(t.w11, d[h].w, fmaf(t.w10, cc[h].w, fmaf(t.w01, b[h].w, t.w00 * a[h].w)));
  }
}
// colours of a dense (N,Hf,Wf,4) image at the taps, `rgb * scale + shift` (unpreprocess) applied per tap
__device__ __forceinline__ void lean_fetch_rgb(const bmv_raygen_fetch_params& p, int view, const LeanTaps& t, float* f) {
  const float* fr = p.rgb + (int64_t)view * p.rgb_view_stride + t.p00 * 4;
  const float sc = p.rgb_scale, sf = p.rgb_shift;
  const float4 a = ldg4(fr), b = ldg4(fr + t.dxp * 4), cc = ldg4(fr + t.dyp * 4), d = ldg4(fr + (t.dyp + t.dxp) * 4);
  f[0] = fmaf(t.w11, fmaf(d.x, sc, sf), fmaf(t.w10, fmaf(cc.x, sc, sf), fmaf(t.w01, fmaf(b.x, sc, sf), t.w00 * fmaf(a.x, sc, sf))));
  f[1] = fmaf(t.w11, fmaf(d.y, sc, sf), fmaf(t.w10, fmaf(cc.y, sc, sf), fmaf(t.w01, fmaf(b.y, sc, sf), t.w00 * fmaf(a.y, sc, sf))));
  f[2] = fmaf(t.w11, fmaf(d.z, sc, sf), fmaf(t.w10, fmaf(cc.z, sc, sf), fmaf(t.w01, fmaf(b.z, sc, sf), t.w00 * fmaf(a.z, sc, sf))));
}
// direction features of a view: normalised difference of the target and source unit directions, and their dot product
__device__ __forceinline__ void lean_dir_feat(const LeanCam& cam, const LeanPoint& q, const float3& tt, float* d) {
  const float4 cc = cam.c;
  float sx = sub_rn(q.x, cc.x), sy = sub_rn(q.y, cc.y), sz = sub_rn(q.zz, cc.z);
  const float n = sqrt_approx(sx * sx + sy * sy + sz * sz) + 1e-6f;
  const float rinv = rcp_approx(n);
  sx *= rinv; sy *= rinv; sz *= rinv;
  const float ex = sub_rn(tt.x, sx), ey = sub_rn(tt.y, sy), ez = sub_rn(tt.z, sz);
  const float en = fmaxf(sqrt_approx(ex * ex + ey * ey + ez * ez), 1e-6f);
  const float re = rcp_approx(en);
  d[0] = ex * re; d[1] = ey * re; d[2] = ez * re;
  d[3] = tt.x * sx + tt.y * sy + tt.z * sz;
}

// RN(cnt / 3) for cnt = 0..3 (the reference's `m /= V`): constants instead of an IEEE division
__device__ __forceinline__ float lean_vis_score3(int cnt) {
  return cnt == 0 ? 0.f : (cnt == 1 ? 0.333333343267440796f : (cnt == 2 ? 0.666666686534881592f : 1.f));
}

}  // namespace bmv
