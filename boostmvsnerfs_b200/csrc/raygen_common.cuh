// Device code shared by the stand-alone K3 (raygen_fetch.cu) and the fused render kernel
// (render_fused.cu): camera block, visibility test, ATen-compatible bilinear/trilinear taps,
// depth-guided ray set-up and sampling.  Reference semantics: lib/networks/enerf/utils.py
// :392-443 (build_rays, sample_along_depth), :458-460 (get_vox_feat), :490-520 (mask_viewport),
// :753-786 (get_img_feat).
#pragma once
#include "bmv_internal.cuh"

namespace bmv {

struct ViewCam {
  float E[12];   // rows 0..2 of world->cam
  float K[9];    // full-resolution intrinsics
  float c[3];    // camera centre in world space
};

// ------------------------------------------------------------------ 3-D visibility of one point
// c = xyz @ R^T (bmm), c += T, q = c @ K^T (bmm), q.xy /= q.z, q.xy /= (W-1,H-1); inside test.
// Every step is a separately rounded fp32 op in the reference; reproduced 1:1 so the integer
// count is bit-exact for identical xyz.
__device__ __forceinline__ bool point_visible(const ViewCam& cam, float x, float y, float z, float isx, float isy) {
  float cx = add_rn(dot3_gemm(x, y, z, cam.E[0], cam.E[1], cam.E[2]), cam.E[3]);
  float cy = add_rn(dot3_gemm(x, y, z, cam.E[4], cam.E[5], cam.E[6]), cam.E[7]);
  float cz = add_rn(dot3_gemm(x, y, z, cam.E[8], cam.E[9], cam.E[10]), cam.E[11]);
  float qx = dot3_gemm(cx, cy, cz, cam.K[0], cam.K[1], cam.K[2]);
  float qy = dot3_gemm(cx, cy, cz, cam.K[3], cam.K[4], cam.K[5]);
  float qz = dot3_gemm(cx, cy, cz, cam.K[6], cam.K[7], cam.K[8]);
  float u = div_rn(div_rn(qx, qz), isx);
  float v = div_rn(div_rn(qy, qz), isy);
  return (u >= 0.f) && (u <= 1.f) && (v >= 0.f) && (v <= 1.f) && (qz > 0.f);
}

__device__ __forceinline__ void load_cam(ViewCam* dst, const float* exts, const float* ixts, const float* centers,
                                         int view, int lane) {
  // 24 values per view, one per thread
  if (lane < 12) dst->E[lane] = exts[view * 16 + lane];
  else if (lane < 21) dst->K[lane - 12] = ixts[view * 9 + (lane - 12)];
  else if (lane < 24) dst->c[lane - 21] = centers ? centers[view * 3 + (lane - 21)] : 0.f;
}

struct Tap2 { int o00, o01, o10, o11; float w00, w01, w10, w11; };

// bilinear, padding_mode='border', align_corners=True, ATen order of operations
__device__ __forceinline__ Tap2 border_taps(float gx, float gy, int H, int W, int64_t ys, int64_t xs) {
  float ix = unnormalize_ac(gx, W), iy = unnormalize_ac(gy, H);
  ix = fminf((float)(W - 1), fmaxf(ix, 0.f));
  iy = fminf((float)(H - 1), fmaxf(iy, 0.f));
  Tap2 t;
  float x0 = floorf(ix), y0 = floorf(iy), x1 = x0 + 1.f, y1 = y0 + 1.f;
  float wx1 = ix - x0, wx0 = x1 - ix, wy1 = iy - y0, wy0 = y1 - iy;
  bool vx1 = x1 <= (float)(W - 1), vy1 = y1 <= (float)(H - 1);
  int ix0 = (int)x0, iy0 = (int)y0, ix1 = vx1 ? ix0 + 1 : ix0, iy1 = vy1 ? iy0 + 1 : iy0;
  t.o00 = (int)(iy0 * ys + ix0 * xs); t.w00 = wx0 * wy0;
  t.o01 = (int)(iy0 * ys + ix1 * xs); t.w01 = vx1 ? wx1 * wy0 : 0.f;
  t.o10 = (int)(iy1 * ys + ix0 * xs); t.w10 = vy1 ? wx0 * wy1 : 0.f;
  t.o11 = (int)(iy1 * ys + ix1 * xs); t.w11 = (vx1 && vy1) ? wx1 * wy1 : 0.f;
  return t;
}
__device__ __forceinline__ float tap2_fetch(const float* __restrict__ f, const Tap2& t) {
  float v = t.w00 * __ldg(f + t.o00);
  v = fmaf(t.w01, __ldg(f + t.o01), v);
  v = fmaf(t.w10, __ldg(f + t.o10), v);
  v = fmaf(t.w11, __ldg(f + t.o11), v);
  return v;
}


// ------------------------------------------------------------------ per-ray set-up (build_rays)
struct RaySetup {
  float ox, oy, oz, dx, dy, dz, fx, fy;   // origin, direction, pixel (as float, like rays[:,6:8])
  float rn, rf, nf0, nf1;                 // ray interval, volume interval
};

__device__ __forceinline__ RaySetup ray_setup(const bmv_raygen_fetch_params& p, int64_t li) {
  RaySetup r;
  float4 ra, rb;
  if (p.rays12_in) {
    const float4* q = reinterpret_cast<const float4*>(p.rays12_in + (p.ray_begin + li) * 12);
    ra = __ldg(q); rb = __ldg(q + 1);
    const float4 rc = __ldg(q + 2);
    r.rn = rc.x; r.rf = rc.y; r.nf0 = rc.z; r.nf1 = rc.w;
  } else {
    const int64_t ri = p.ray_begin + li;
    if (p.rays) {
      ra = __ldg(reinterpret_cast<const float4*>(p.rays + ri * 8));
      rb = __ldg(reinterpret_cast<const float4*>(p.rays + ri * 8 + 4));
    } else {                                              // generate the ray of pixel (ri % W, ri / W)
      const double* G = p.ray_gen;
      const int gx = (int)(ri % p.W), gy = (int)(ri / p.W);
      const double dxp = (double)gx, dyp = (double)gy;
      const double d0 = __fma_rn(dyp, __ldg(G + 6), __dmul_rn(dxp, __ldg(G + 3))) + __ldg(G + 9);
      const double d1 = __fma_rn(dyp, __ldg(G + 7), __dmul_rn(dxp, __ldg(G + 4))) + __ldg(G + 10);
      const double d2 = __fma_rn(dyp, __ldg(G + 8), __dmul_rn(dxp, __ldg(G + 5))) + __ldg(G + 11);
      ra = make_float4((float)__ldg(G), (float)__ldg(G + 1), (float)__ldg(G + 2), (float)d0);
      rb = make_float4((float)d1, (float)d2, (float)gx, (float)gy);
    }
    int px = (int)rb.z, py = (int)rb.w;                 // .long(): truncation toward zero
    px = min(max(px, 0), p.W - 1);
    py = min(max(py, 0), p.H - 1);
    // upsample the per-pixel depth interval to the render grid and clamp it
    const UpCoord uy = up_coord(py, p.hv, p.H), ux = up_coord(px, p.wv, p.W);
    const int hwv = p.hv * p.wv;
    const float dep = up_sample(p.depth, p.wv, uy, ux);
    const float sd = up_sample(p.std, p.wv, uy, ux);
    r.nf0 = up_sample(p.near_far, p.wv, uy, ux);
    r.nf1 = up_sample(p.near_far + hwv, p.wv, uy, ux);
    if (p.depth_inv) {
      r.rn = add_rn(dep, sd); r.rf = sub_rn(dep, sd);
      r.rn = r.rn > r.nf0 ? r.nf0 : r.rn;
      r.rf = r.rf < r.nf1 ? r.nf1 : r.rf;
    } else {
      r.rn = sub_rn(dep, sd); r.rf = add_rn(dep, sd);
      r.rn = r.rn < r.nf0 ? r.nf0 : r.rn;
      r.rf = r.rf > r.nf1 ? r.nf1 : r.rf;
    }
  }
  r.ox = ra.x; r.oy = ra.y; r.oz = ra.z; r.dx = ra.w; r.dy = rb.x; r.dz = rb.y; r.fx = rb.z; r.fy = rb.w;
  return r;
}

// ------------------------------------------------------------------ sample_along_depth, sample s
struct SamplePoint { float z, x, y, zz, dn; };
__device__ __forceinline__ SamplePoint sample_point(const bmv_raygen_fetch_params& p, const RaySetup& r, int s) {
  SamplePoint q;
  const float t = (p.S == 1) ? 0.5f : __ldg(p.t + s);
  q.z = add_rn(r.rn, mul_rn(sub_rn(r.rf, r.rn), t));
  if (p.depth_inv) {
    const float iz = div_rn(1.f, fmaxf(q.z, 1e-6f));
    q.x = add_rn(r.ox, mul_rn(r.dx, iz)); q.y = add_rn(r.oy, mul_rn(r.dy, iz)); q.zz = add_rn(r.oz, mul_rn(r.dz, iz));
    q.dn = div_rn(sub_rn(r.nf0, q.z), fmaxf(sub_rn(r.nf0, r.nf1), 1e-6f));
  } else {
    q.x = add_rn(r.ox, mul_rn(r.dx, q.z)); q.y = add_rn(r.oy, mul_rn(r.dy, q.z)); q.zz = add_rn(r.oz, mul_rn(r.dz, q.z));
    q.dn = div_rn(sub_rn(q.z, r.nf0), fmaxf(sub_rn(r.nf1, r.nf0), 1e-6f));
  }
  return q;
}

// ------------------------------------------------------------------ register-resident gather
// Same arithmetic as fetch_sample() in raygen_fetch.cu, but for compile-time channel counts and with
// the results left in registers: vox[8], f[v] = [CF image-feature channels, rgb(3), dir(4)], and the
// visibility count.  Feeds nerf_mlp_eval() directly.
// rgb strides with the all-zero default (planar contiguous)
struct RgbStrides { int64_t c, y, x; };
__device__ __forceinline__ RgbStrides rgb_strides(const bmv_raygen_fetch_params& p) {
  RgbStrides r;
  if (p.rgb_x_stride == 0 && p.rgb_y_stride == 0 && p.rgb_c_stride == 0) { r.c = (int64_t)p.Hf * p.Wf; r.y = p.Wf; r.x = 1; }
  else { r.c = p.rgb_c_stride; r.y = p.rgb_y_stride; r.x = p.rgb_x_stride; }
  return r;
}
// Host-side test for the VEC instantiation of gather_sample_regs: 8-channel channels-last volume and image
// features with 16-byte aligned voxels / pixels, and a 4-float-per-pixel channels-last rgb image.
inline bool gather_vec_ok(const bmv_raygen_fetch_params& p) {
  auto a16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  auto m4 = [](int64_t v) { return v % 4 == 0; };
  return p.Cv == 8 && p.Cf == 8 && p.vol_c_stride == 1 && p.imf_c_stride == 1 && a16(p.volume) && a16(p.im_feat) &&
         a16(p.rgb) && m4(p.vol_d_stride) && m4(p.vol_y_stride) && m4(p.vol_x_stride) && m4(p.imf_view_stride) &&
         m4(p.imf_y_stride) && m4(p.imf_x_stride) && p.rgb_c_stride == 1 && p.rgb_x_stride == 4 && m4(p.rgb_y_stride) &&
         m4(p.rgb_view_stride);
}

__device__ __forceinline__ float4 ldg4(const float* q) { return __ldg(reinterpret_cast<const float4*>(q)); }

// Division for quantities that only steer interpolation / direction features (tolerance 1e-4): the VEC (fast
// path) instantiation uses MUFU.RCP + multiply (<= 2 ulp) instead of the ~15-instruction IEEE sequence — 33
// divisions per sample.  Everything that decides a visibility count or a sample depth stays IEEE (div_rn).
template <bool FAST>
__device__ __forceinline__ float gdiv(float a, float b) { return FAST ? __fdividef(a, b) : div_rn(a, b); }

// VEC = true: the layouts of gather_vec_ok(); every tap is fetched with 16-byte loads (52 loads per sample
// instead of 196 scalar ones).  The arithmetic — order of the fmaf chains per channel — is identical.
template <int CF, int V, bool VEC = false>
__device__ __forceinline__ int gather_sample_regs(const bmv_raygen_fetch_params& p, const ViewCam* cams,
                                                  const int* views, const float* tar_c, float x, float y, float zz,
                                                  float gxv, float gyv, float dn, float (&vox)[8],
                                                  float (&f)[V][CF + 7]) {
  const float isx = (float)(p.W - 1), isy = (float)(p.H - 1);
  {
    const float gz = sub_rn(mul_rn(dn, 2.f), 1.f);
    const float ix = unnormalize_ac(gxv, p.wv), iy = unnormalize_ac(gyv, p.hv), iz = unnormalize_ac(gz, p.Dv);
#pragma unroll
    for (int c = 0; c < 8; ++c) vox[c] = 0.f;
    if (coord_ok(ix) && coord_ok(iy) && coord_ok(iz)) {
      const float x0 = floorf(ix), y0 = floorf(iy), z0 = floorf(iz);
      const float fx1 = ix - x0, fy1 = iy - y0, fz1 = iz - z0;
      const float fx0 = (x0 + 1.f) - ix, fy0 = (y0 + 1.f) - iy, fz0 = (z0 + 1.f) - iz;
#pragma unroll
      for (int corner = 0; corner < 8; ++corner) {
        const int bx = corner & 1, by = (corner >> 1) & 1, bz = corner >> 2;
        const float cxf = x0 + bx, cyf = y0 + by, czf = z0 + bz;
        const bool ok = cxf >= 0.f && cxf <= (float)(p.wv - 1) && cyf >= 0.f && cyf <= (float)(p.hv - 1) &&
                        czf >= 0.f && czf <= (float)(p.Dv - 1);
        if (!ok) continue;
        const float w = (bx ? fx1 : fx0) * (by ? fy1 : fy0) * (bz ? fz1 : fz0);
        const float* src = p.volume + (int64_t)czf * p.vol_d_stride + (int64_t)cyf * p.vol_y_stride +
                           (int64_t)cxf * p.vol_x_stride;
        if (VEC) {
          const float4 a = ldg4(src), b = ldg4(src + 4);
          vox[0] = fmaf(w, a.x, vox[0]); vox[1] = fmaf(w, a.y, vox[1]); vox[2] = fmaf(w, a.z, vox[2]); vox[3] = fmaf(w, a.w, vox[3]);
          vox[4] = fmaf(w, b.x, vox[4]); vox[5] = fmaf(w, b.y, vox[5]); vox[6] = fmaf(w, b.z, vox[6]); vox[7] = fmaf(w, b.w, vox[7]);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c) vox[c] = fmaf(w, __ldg(src + (int64_t)c * p.vol_c_stride), vox[c]);
        }
      }
    }
  }
  int cnt = 0;
  float ttx = sub_rn(x, tar_c[0]), tty = sub_rn(y, tar_c[1]), ttz = sub_rn(zz, tar_c[2]);
  {
    const float n = sqrtf(ttx * ttx + tty * tty + ttz * ttz) + 1e-6f;
    ttx = gdiv<VEC>(ttx, n); tty = gdiv<VEC>(tty, n); ttz = gdiv<VEC>(ttz, n);
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const ViewCam& cam = cams[v];
    cnt += point_visible(cam, x, y, zz, isx, isy) ? 1 : 0;
    const float cx = dot4_gemm(x, y, zz, 1.f, cam.E[0], cam.E[1], cam.E[2], cam.E[3]);
    const float cy = dot4_gemm(x, y, zz, 1.f, cam.E[4], cam.E[5], cam.E[6], cam.E[7]);
    const float cz = dot4_gemm(x, y, zz, 1.f, cam.E[8], cam.E[9], cam.E[10], cam.E[11]);
    const float rs = p.render_scale;
    const float qx = dot3_gemm(cx, cy, cz, mul_rn(cam.K[0], rs), mul_rn(cam.K[1], rs), mul_rn(cam.K[2], rs));
    const float qy = dot3_gemm(cx, cy, cz, mul_rn(cam.K[3], rs), mul_rn(cam.K[4], rs), mul_rn(cam.K[5], rs));
    const float qz = dot3_gemm(cx, cy, cz, cam.K[6], cam.K[7], cam.K[8]);
    const float qzc = (qz != qz) ? qz : fmaxf(qz, 1e-6f);
    float gx = gdiv<VEC>(gdiv<VEC>(qx, qzc), (float)(p.Wf - 1));
    float gy = gdiv<VEC>(gdiv<VEC>(qy, qzc), (float)(p.Hf - 1));
    gx = sub_rn(mul_rn(gx, 2.f), 1.f);
    gy = sub_rn(mul_rn(gy, 2.f), 1.f);
    const int view = views[v];
    if (VEC) {
      const Tap2 tp = border_taps(gx, gy, p.Hf, p.Wf, p.imf_y_stride, p.imf_x_stride);
      const float* fm = p.im_feat + (int64_t)view * p.imf_view_stride;
#pragma unroll
      for (int h = 0; h < CF / 4; ++h) {
        const float4 a = ldg4(fm + tp.o00 + 4 * h), b = ldg4(fm + tp.o01 + 4 * h);
        const float4 c = ldg4(fm + tp.o10 + 4 * h), d = ldg4(fm + tp.o11 + 4 * h);
        f[v][4 * h + 0] = fmaf(tp.w11, d.x, fmaf(tp.w10, c.x, fmaf(tp.w01, b.x, tp.w00 * a.x)));
        f[v][4 * h + 1] = fmaf(tp.w11, d.y, fmaf(tp.w10, c.y, fmaf(tp.w01, b.y, tp.w00 * a.y)));
        f[v][4 * h + 2] = fmaf(tp.w11, d.z, fmaf(tp.w10, c.z, fmaf(tp.w01, b.z, tp.w00 * a.z)));
        f[v][4 * h + 3] = fmaf(tp.w11, d.w, fmaf(tp.w10, c.w, fmaf(tp.w01, b.w, tp.w00 * a.w)));
      }
      // rgb: (N,Hf,Wf,4) channels-last, same taps at 4 floats per pixel
      const Tap2 tr = border_taps(gx, gy, p.Hf, p.Wf, p.rgb_y_stride, 4);
      const float* fr = p.rgb + (int64_t)view * p.rgb_view_stride;
      const float4 a = ldg4(fr + tr.o00), b = ldg4(fr + tr.o01), c = ldg4(fr + tr.o10), d = ldg4(fr + tr.o11);
      const float sc = p.rgb_scale, sf = p.rgb_shift;
      f[v][CF + 0] = fmaf(tr.w11, fmaf(d.x, sc, sf), fmaf(tr.w10, fmaf(c.x, sc, sf), fmaf(tr.w01, fmaf(b.x, sc, sf), tr.w00 * fmaf(a.x, sc, sf))));
      f[v][CF + 1] = fmaf(tr.w11, fmaf(d.y, sc, sf), fmaf(tr.w10, fmaf(c.y, sc, sf), fmaf(tr.w01, fmaf(b.y, sc, sf), tr.w00 * fmaf(a.y, sc, sf))));
      f[v][CF + 2] = fmaf(tr.w11, fmaf(d.z, sc, sf), fmaf(tr.w10, fmaf(c.z, sc, sf), fmaf(tr.w01, fmaf(b.z, sc, sf), tr.w00 * fmaf(a.z, sc, sf))));
    } else {
      {
        const Tap2 tp = border_taps(gx, gy, p.Hf, p.Wf, p.imf_y_stride, p.imf_x_stride);
        const float* fm = p.im_feat + (int64_t)view * p.imf_view_stride;
#pragma unroll
        for (int c = 0; c < CF; ++c) f[v][c] = tap2_fetch(fm + (int64_t)c * p.imf_c_stride, tp);
      }
      const RgbStrides rs3 = rgb_strides(p);
      const Tap2 tp = border_taps(gx, gy, p.Hf, p.Wf, rs3.y, rs3.x);
      const float* fm = p.rgb + (int64_t)view * p.rgb_view_stride;
      const int64_t plane = rs3.c;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* fc = fm + c * plane;
        float val = tp.w00 * fmaf(__ldg(fc + tp.o00), p.rgb_scale, p.rgb_shift);
        val = fmaf(tp.w01, fmaf(__ldg(fc + tp.o01), p.rgb_scale, p.rgb_shift), val);
        val = fmaf(tp.w10, fmaf(__ldg(fc + tp.o10), p.rgb_scale, p.rgb_shift), val);
        val = fmaf(tp.w11, fmaf(__ldg(fc + tp.o11), p.rgb_scale, p.rgb_shift), val);
        f[v][CF + c] = val;
      }
    }
    float sx = sub_rn(x, cam.c[0]), sy = sub_rn(y, cam.c[1]), sz = sub_rn(zz, cam.c[2]);
    const float n = sqrtf(sx * sx + sy * sy + sz * sz) + 1e-6f;
    sx = gdiv<VEC>(sx, n); sy = gdiv<VEC>(sy, n); sz = gdiv<VEC>(sz, n);
    const float ex = sub_rn(ttx, sx), ey = sub_rn(tty, sy), ez = sub_rn(ttz, sz);
    const float en = fmaxf(sqrtf(ex * ex + ey * ey + ez * ez), 1e-6f);
    f[v][CF + 3] = gdiv<VEC>(ex, en);
    f[v][CF + 4] = gdiv<VEC>(ey, en);
    f[v][CF + 5] = gdiv<VEC>(ez, en);
    f[v][CF + 6] = ttx * sx + tty * sy + ttz * sz;
  }
  return cnt;
}

}  // namespace bmv
