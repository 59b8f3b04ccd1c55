// Argument checks shared by the multi-chain render launchers (bmv_render_rays_multi, bmv_render_rays_multi_umma):
// both read the same bmv_render_multi_params and make the same layout assumptions (lean_gather.cuh).
#pragma once
#include "bmv_internal.cuh"

namespace bmv {

inline int render_multi_validate(const bmv_render_multi_params* mp, const char* who) {
  BMV_REQUIRE(mp != nullptr, BMV_ERR_INVALID_ARGUMENT, "%s: null params", who);
  const bmv_raygen_fetch_params* p = &mp->g;
  BMV_REQUIRE(p->n_rays >= 0 && p->ray_begin >= 0, BMV_ERR_INVALID_ARGUMENT, "%s: bad ray range", who);
  BMV_REQUIRE(mp->K >= 1 && mp->K <= BMV_MAX_VOLUMES && mp->n_views >= 1 && mp->n_views <= BMV_MAX_VIEWS,
              BMV_ERR_INVALID_ARGUMENT, "%s: K must be 1..%d and n_views 1..%d", who, BMV_MAX_VOLUMES, BMV_MAX_VIEWS);
  if (p->n_rays == 0) return BMV_OK;                       // nothing to validate against
  BMV_REQUIRE(mp->depth && mp->std && mp->near_far && mp->volume && (p->rays || p->ray_gen) && p->im_feat && p->rgb &&
                  p->src_exts && p->src_ixts && p->src_centers && p->tar_center && mp->mlp_weights && mp->raw,
              BMV_ERR_INVALID_ARGUMENT, "%s: null device pointer", who);
  BMV_REQUIRE(((uintptr_t)mp->mlp_weights & 15) == 0 && ((uintptr_t)mp->raw & 15) == 0 && ((uintptr_t)p->rgb & 15) == 0 &&
                  (!p->rays || ((uintptr_t)p->rays & 15) == 0),
              BMV_ERR_INVALID_ARGUMENT, "%s: pointers must be 16-byte aligned", who);
  // 8-channel texels are fetched with one 256-bit load each
  BMV_REQUIRE(((uintptr_t)mp->volume & 31) == 0 && ((uintptr_t)p->im_feat & 31) == 0 && mp->vol_k_stride % 8 == 0 &&
                  p->vol_d_stride % 8 == 0 && p->imf_view_stride % 8 == 0,
              BMV_ERR_INVALID_ARGUMENT, "%s: volumes / feature maps must be 32-byte aligned (base and every chain / plane / view)", who);
  BMV_REQUIRE(p->S >= 1 && (p->S == 1 || p->t), BMV_ERR_INVALID_ARGUMENT, "%s: bad S / t", who);
  BMV_REQUIRE(p->H >= 2 && p->W >= 2 && p->hv >= 1 && p->wv >= 1 && p->Hf >= 2 && p->Wf >= 2 && p->Dv >= 1,
              BMV_ERR_INVALID_ARGUMENT, "%s: bad grid size", who);
  BMV_REQUIRE(p->Cv == 8 && p->Cf == 8 && p->V == 3, BMV_ERR_UNSUPPORTED_SHAPE,
              "%s: (Cv=%d, Cf=%d, V=%d) not instantiated (8, 8, 3)", who, p->Cv, p->Cf, p->V);
  // dense channels-last layouts, 32-bit offsets
  BMV_REQUIRE(p->vol_c_stride == 1 && p->vol_x_stride == 8 && p->vol_y_stride == (int64_t)p->wv * 8 &&
                  p->vol_d_stride % 4 == 0 && p->vol_d_stride > 0 && p->vol_d_stride <= (int64_t)p->hv * p->wv * 8 &&
                  mp->vol_k_stride % 4 == 0,
              BMV_ERR_UNSUPPORTED_SHAPE, "%s: the volumes must be (D,rows,w,8) channels-last with dense rows", who);
  BMV_REQUIRE(mp->vol_row0 >= 0 && mp->vol_row0 < p->hv && mp->map_row0 >= 0 && mp->map_row0 < p->hv && mp->nf_plane_stride >= 0,
              BMV_ERR_INVALID_ARGUMENT, "%s: bad slab rows", who);
  BMV_REQUIRE(p->imf_c_stride == 1 && p->imf_x_stride == 8 && p->imf_y_stride == (int64_t)p->Wf * 8 && p->imf_view_stride % 4 == 0,
              BMV_ERR_UNSUPPORTED_SHAPE, "%s: im_feat must be dense (N,Hf,Wf,8) channels-last", who);
  BMV_REQUIRE(p->rgb_c_stride == 1 && p->rgb_x_stride == 4 && p->rgb_y_stride == (int64_t)p->Wf * 4 && p->rgb_view_stride % 4 == 0,
              BMV_ERR_UNSUPPORTED_SHAPE, "%s: rgb must be dense (N,Hf,Wf,4)", who);
  BMV_REQUIRE((int64_t)p->Dv * p->vol_d_stride < (1ll << 31) && (int64_t)p->Hf * p->Wf * 8 < (1ll << 31) &&
                  p->n_rays * p->S < (1ll << 31) && p->ray_begin + p->n_rays < (1ll << 31),
              BMV_ERR_UNSUPPORTED_SHAPE, "%s: tensors too large for 32-bit offsets", who);
  for (int i = 0; !mp->views && i < mp->K * 3; ++i)
    BMV_REQUIRE(mp->views_host[i] >= 0 && mp->views_host[i] < mp->n_views, BMV_ERR_INVALID_ARGUMENT,
                "%s: view id %d out of range", who, mp->views_host[i]);
  return BMV_OK;
}

}  // namespace bmv
