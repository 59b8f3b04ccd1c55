// K3: depth-guided ray generation + sampling + trilinear/bilinear feature fetch + 3-D visibility.
// Reference semantics (lib/networks/enerf/utils.py): build_rays :392-422, sample_along_depth
// :424-443, get_vox_feat :458-460, get_ndc_coords/mask_viewport :490-520, unpreprocess :669-676,
// get_img_feat :753-786; glue lib/networks/boost_enerf/network.py:123-149.
#include "raygen_common.cuh"

namespace bmv {

// Everything that happens to ONE sample once its world position is known: trilinear volume fetch,
// per-view colour/feature fetch + direction features, visibility count.
template <int MAXV>
__device__ __forceinline__ void fetch_sample(const bmv_raygen_fetch_params& p, const ViewCam* cams, const int* views,
                                             const float* tar_c, int64_t si, float x, float y, float zz,
                                             float gxv, float gyv, float dn) {
  const int V = p.V, Cf = p.Cf, Cv = p.Cv;
  const int row = Cf + 7;                               // per-view feature row: Cf + rgb3 + dir4
  const float isx = (float)(p.W - 1), isy = (float)(p.H - 1);
  // ---- get_vox_feat: trilinear, zeros padding
  if (p.vox_feat) {
    const float gz = sub_rn(mul_rn(dn, 2.f), 1.f);
    const float ix = unnormalize_ac(gxv, p.wv), iy = unnormalize_ac(gyv, p.hv), iz = unnormalize_ac(gz, p.Dv);
    float acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) acc[c] = 0.f;
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
#pragma unroll
        for (int c = 0; c < 8; ++c)
          if (c < Cv) acc[c] = fmaf(w, __ldg(src + (int64_t)c * p.vol_c_stride), acc[c]);
      }
    }
    float* o = p.vox_feat + si * Cv;
    if (Cv == 8) {
      reinterpret_cast<float4*>(o)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      reinterpret_cast<float4*>(o)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        if (c < Cv) o[c] = acc[c];
    }
  }
  if (!p.img_feat && !p.vis_count && !p.vis_mask) return;

  // ---- per source view: colour/feature fetch, direction features, visibility
  int cnt = 0;
  float ttx = sub_rn(x, tar_c[0]), tty = sub_rn(y, tar_c[1]), ttz = sub_rn(zz, tar_c[2]);
  {
    const float n = sqrtf(ttx * ttx + tty * tty + ttz * ttz) + 1e-6f;
    ttx = div_rn(ttx, n); tty = div_rn(tty, n); ttz = div_rn(ttz, n);
  }
  for (int v = 0; v < V; ++v) {
    const ViewCam& cam = cams[v];
    cnt += point_visible(cam, x, y, zz, isx, isy) ? 1 : 0;
    if (!p.img_feat) continue;
    // xyz1 @ ext^T (4-term chain), then @ (K*render_scale)^T
    const float cx = dot4_gemm(x, y, zz, 1.f, cam.E[0], cam.E[1], cam.E[2], cam.E[3]);
    const float cy = dot4_gemm(x, y, zz, 1.f, cam.E[4], cam.E[5], cam.E[6], cam.E[7]);
    const float cz = dot4_gemm(x, y, zz, 1.f, cam.E[8], cam.E[9], cam.E[10], cam.E[11]);
    const float rs = p.render_scale;
    const float qx = dot3_gemm(cx, cy, cz, mul_rn(cam.K[0], rs), mul_rn(cam.K[1], rs), mul_rn(cam.K[2], rs));
    const float qy = dot3_gemm(cx, cy, cz, mul_rn(cam.K[3], rs), mul_rn(cam.K[4], rs), mul_rn(cam.K[5], rs));
    const float qz = dot3_gemm(cx, cy, cz, cam.K[6], cam.K[7], cam.K[8]);
    const float qzc = (qz != qz) ? qz : fmaxf(qz, 1e-6f);
    float gx = div_rn(div_rn(qx, qzc), (float)(p.Wf - 1));
    float gy = div_rn(div_rn(qy, qzc), (float)(p.Hf - 1));
    gx = sub_rn(mul_rn(gx, 2.f), 1.f);
    gy = sub_rn(mul_rn(gy, 2.f), 1.f);
    float* o = p.img_feat + (si * V + v) * row;
    const int view = views[v];
    {
      const Tap2 tp = border_taps(gx, gy, p.Hf, p.Wf, p.imf_y_stride, p.imf_x_stride);
      const float* f = p.im_feat + (int64_t)view * p.imf_view_stride;
      for (int c = 0; c < Cf; ++c) o[c] = tap2_fetch(f + (int64_t)c * p.imf_c_stride, tp);
    }
    {
      const RgbStrides rs3 = rgb_strides(p);
      const Tap2 tp = border_taps(gx, gy, p.Hf, p.Wf, rs3.y, rs3.x);
      const float* f = p.rgb + (int64_t)view * p.rgb_view_stride;
      const int64_t plane = rs3.c;
      // colour = bilinear(img*scale+shift); img*0.5+0.5 is exact as an FMA (0.5 is a power of two)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float* fc = f + c * plane;
        float val = tp.w00 * fmaf(__ldg(fc + tp.o00), p.rgb_scale, p.rgb_shift);
        val = fmaf(tp.w01, fmaf(__ldg(fc + tp.o01), p.rgb_scale, p.rgb_shift), val);
        val = fmaf(tp.w10, fmaf(__ldg(fc + tp.o10), p.rgb_scale, p.rgb_shift), val);
        val = fmaf(tp.w11, fmaf(__ldg(fc + tp.o11), p.rgb_scale, p.rgb_shift), val);
        o[Cf + c] = val;
      }
    }
    float sx = sub_rn(x, cam.c[0]), sy = sub_rn(y, cam.c[1]), sz = sub_rn(zz, cam.c[2]);
    const float n = sqrtf(sx * sx + sy * sy + sz * sz) + 1e-6f;
    sx = div_rn(sx, n); sy = div_rn(sy, n); sz = div_rn(sz, n);
    const float ex = sub_rn(ttx, sx), ey = sub_rn(tty, sy), ez = sub_rn(ttz, sz);
    const float en = fmaxf(sqrtf(ex * ex + ey * ey + ez * ez), 1e-6f);
    o[Cf + 3] = div_rn(ex, en);
    o[Cf + 4] = div_rn(ey, en);
    o[Cf + 5] = div_rn(ez, en);
    o[Cf + 6] = ttx * sx + tty * sy + ttz * sz;
  }
  if (p.vis_count) p.vis_count[si] = cnt;
  if (p.vis_mask) p.vis_mask[si] = div_rn((float)cnt, (float)V);
}

// v1: one thread per ray, samples handled sequentially.  Three input modes:
//   fused      : depth/std/near_far maps + rays(R,8)            (the per-frame hot path)
//   rays12_in  : rays already carry their interval (R,12)       (function-level sample_along_depth)
//   xyz_in     : explicit points (+ uvd_in for the volume fetch) (function-level get_vox_feat/get_img_feat)
template <int MAXV>
__global__ void __launch_bounds__(128) raygen_fetch_kernel(bmv_raygen_fetch_params p) {
  __shared__ ViewCam cams[MAXV];
  __shared__ float s_tar_c[3];
  __shared__ int s_view[MAXV];
  if (threadIdx.x < MAXV) s_view[threadIdx.x] = threadIdx.x < p.V ? p.view[threadIdx.x] : 0;
  __syncthreads();
  for (int v = 0; v < p.V; ++v) load_cam(&cams[v], p.src_exts, p.src_ixts, p.src_centers, s_view[v], threadIdx.x);
  if (threadIdx.x < 3) s_tar_c[threadIdx.x] = p.tar_center ? p.tar_center[threadIdx.x] : 0.f;
  __syncthreads();
  const int64_t li = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // local ray index
  if (li >= p.n_rays) return;

  if (p.xyz_in) {   // ---- pointwise mode
    const float x = __ldg(p.xyz_in + li * 3), y = __ldg(p.xyz_in + li * 3 + 1), z = __ldg(p.xyz_in + li * 3 + 2);
    float gxv = 0.f, gyv = 0.f, dn = 0.f;
    if (p.uvd_in) {
      gxv = sub_rn(mul_rn(__ldg(p.uvd_in + li * 3), 2.f), 1.f);
      gyv = sub_rn(mul_rn(__ldg(p.uvd_in + li * 3 + 1), 2.f), 1.f);
      dn = __ldg(p.uvd_in + li * 3 + 2);
    }
    fetch_sample<MAXV>(p, cams, s_view, s_tar_c, li, x, y, z, gxv, gyv, dn);
    return;
  }

  const RaySetup r = ray_setup(p, li);
  if (p.rays12) {
    float4* o = reinterpret_cast<float4*>(p.rays12 + li * 12);
    o[0] = make_float4(r.ox, r.oy, r.oz, r.dx);
    o[1] = make_float4(r.dy, r.dz, r.fx, r.fy);
    o[2] = make_float4(r.rn, r.rf, r.nf0, r.nf1);
  }
  const int S = p.S;
  const float un = div_rn(r.fx, (float)(p.W - 1)), vn = div_rn(r.fy, (float)(p.H - 1));
  const float gxv = sub_rn(mul_rn(un, 2.f), 1.f), gyv = sub_rn(mul_rn(vn, 2.f), 1.f);
  const bool need_fetch = p.vox_feat || p.img_feat || p.vis_count || p.vis_mask;
  if (!(need_fetch || p.z_vals || p.xyz || p.uvd)) return;
  for (int s = 0; s < S; ++s) {
    const SamplePoint q = sample_point(p, r, s);
    const int64_t si = li * S + s;
    if (p.z_vals) p.z_vals[si] = q.z;
    if (p.xyz) { p.xyz[si * 3] = q.x; p.xyz[si * 3 + 1] = q.y; p.xyz[si * 3 + 2] = q.zz; }
    if (p.uvd) { p.uvd[si * 3] = r.fx; p.uvd[si * 3 + 1] = r.fy; p.uvd[si * 3 + 2] = q.dn; }
    if (need_fetch) fetch_sample<MAXV>(p, cams, s_view, s_tar_c, si, q.x, q.y, q.zz, gxv, gyv, q.dn);
  }
}

template <int MAXV>
__global__ void __launch_bounds__(256) mask_viewport_kernel(bmv_visibility_params p) {
  __shared__ ViewCam cams[MAXV];
  __shared__ int s_view[MAXV];
  if (threadIdx.x < MAXV) s_view[threadIdx.x] = threadIdx.x < p.V ? p.view[threadIdx.x] : 0;
  __syncthreads();
  for (int v = 0; v < p.V; ++v) load_cam(&cams[v], p.src_exts, p.src_ixts, nullptr, s_view[v], threadIdx.x);
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n_pts) return;
  const float x = __ldg(p.xyz + i * 3), y = __ldg(p.xyz + i * 3 + 1), z = __ldg(p.xyz + i * 3 + 2);
  int cnt = 0;
  for (int v = 0; v < p.V; ++v) cnt += point_visible(cams[v], x, y, z, p.inv_scale_x, p.inv_scale_y) ? 1 : 0;
  if (p.vis_count) p.vis_count[i] = cnt;
  if (p.vis_mask) p.vis_mask[i] = div_rn((float)cnt, (float)p.V);
}

}  // namespace bmv

extern "C" BMV_API int bmv_raygen_sample_fetch(const bmv_raygen_fetch_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_raygen_sample_fetch");
  using namespace bmv;
  BMV_REQUIRE(p != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_raygen_sample_fetch: null params");
  BMV_REQUIRE(p->src_exts && p->src_ixts, BMV_ERR_INVALID_ARGUMENT, "bmv_raygen_sample_fetch: null camera pointer");
  if (!p->xyz_in && !p->rays12_in)
    BMV_REQUIRE(p->depth && p->std && p->near_far && (p->rays || p->ray_gen), BMV_ERR_INVALID_ARGUMENT,
                "bmv_raygen_sample_fetch: null input pointer");
  if (p->xyz_in && p->vox_feat)
    BMV_REQUIRE(p->uvd_in != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_raygen_sample_fetch: vox_feat needs uvd_in");
  BMV_REQUIRE(p->S >= 1 && (p->S == 1 || p->t || p->xyz_in), BMV_ERR_INVALID_ARGUMENT,
              "bmv_raygen_sample_fetch: bad S / t");
  BMV_REQUIRE(p->V >= 1 && p->V <= BMV_MAX_VIEWS, BMV_ERR_INVALID_ARGUMENT, "bmv_raygen_sample_fetch: bad V=%d", p->V);
  BMV_REQUIRE(p->H >= 2 && p->W >= 2 && p->hv >= 1 && p->wv >= 1, BMV_ERR_INVALID_ARGUMENT,
              "bmv_raygen_sample_fetch: bad grid size");
  BMV_REQUIRE(p->n_rays >= 0 && p->ray_begin >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_raygen_sample_fetch: bad ray range");
  if (p->vox_feat) {
    BMV_REQUIRE(p->volume && p->Cv >= 1 && p->Cv <= 8 && p->Dv >= 1, BMV_ERR_UNSUPPORTED_SHAPE,
                "bmv_raygen_sample_fetch: volume channels must be 1..8 (got %d)", p->Cv);
  }
  if (p->img_feat) {
    BMV_REQUIRE(p->im_feat && p->rgb && p->src_centers && p->tar_center && p->Cf >= 0 && p->Hf >= 2 && p->Wf >= 2,
                BMV_ERR_INVALID_ARGUMENT, "bmv_raygen_sample_fetch: image feature inputs missing");
  }
  if (p->n_rays == 0) return BMV_OK;
  const int threads = 128;
  const unsigned blocks = (unsigned)ceil_div64(p->n_rays, threads);
  raygen_fetch_kernel<BMV_MAX_VIEWS><<<blocks, threads, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_raygen_sample_fetch");
}

extern "C" BMV_API int bmv_mask_viewport(const bmv_visibility_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_mask_viewport");
  using namespace bmv;
  BMV_REQUIRE(p && p->xyz && p->src_exts && p->src_ixts, BMV_ERR_INVALID_ARGUMENT, "bmv_mask_viewport: null pointer");
  BMV_REQUIRE(p->V >= 1 && p->V <= BMV_MAX_VIEWS, BMV_ERR_INVALID_ARGUMENT, "bmv_mask_viewport: bad V=%d", p->V);
  BMV_REQUIRE(p->n_pts >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_mask_viewport: negative n_pts");
  if (p->n_pts == 0) return BMV_OK;
  const unsigned blocks = (unsigned)ceil_div64(p->n_pts, 256);
  mask_viewport_kernel<BMV_MAX_VIEWS><<<blocks, 256, 0, (cudaStream_t)stream>>>(*p);
  return check_launch("bmv_mask_viewport");
}
