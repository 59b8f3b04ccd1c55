/* ORACLE — TEST INFRASTRUCTURE ONLY (never linked into or called by the product library).
 *
 * Plain-C restatement of the two pieces of the hot path whose results are compared BIT-EXACTLY
 * (visibility counts) or that are pure scalar recurrences (the K-volume blend), written with
 * explicit fmaf()/separately rounded operations so that the exact fp32 operation order the CUDA
 * kernels reproduce is pinned independently of PyTorch:
 *   oracle_visibility_count  — reference lib/networks/enerf/utils.py:490-520 (get_ndc_coords + mask_viewport)
 *   oracle_composite_blend   — reference lib/networks/boost_enerf/network.py:163-170 +
 *                              lib/networks/enerf/utils.py:639-667 (merge_mlp_outputs + raw2outputs_blend)
 * Pinned by tests/test_c_oracle.py against the golden vectors of the UNMODIFIED reference
 * (tests/golden/enerf_ops.npz: mask_wide / blend_*).
 * Build: oracle/build_c.py (gcc -O2 -ffp-contract=off -shared -fPIC) -> oracle/_build/libhotpath_oracle.so
 */
#include <math.h>
#include <stdint.h>

/* torch.bmm inner product for K=3 as an SGEMM micro-kernel accumulates it: k ascending from 0 */
static float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
  float acc = a0 * b0;
  acc = fmaf(a1, b1, acc);
  acc = fmaf(a2, b2, acc);
  return acc;
}

/* xyz (n,3); exts (V,4,4) world->cam; ixts (V,3,3); inv scale (W-1, H-1); count (n) in 0..V */
void oracle_visibility_count(const float* xyz, int64_t n, const float* exts, const float* ixts, int V,
                             float isx, float isy, int32_t* count) {
  for (int64_t i = 0; i < n; ++i) {
    const float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    int c = 0;
    for (int v = 0; v < V; ++v) {
      const float* E = exts + 16 * v;
      const float* K = ixts + 9 * v;
      /* cam = xyz @ R^T (bmm) ; cam += T (add_) */
      const float cx = dot3(x, y, z, E[0], E[1], E[2]) + E[3];
      const float cy = dot3(x, y, z, E[4], E[5], E[6]) + E[7];
      const float cz = dot3(x, y, z, E[8], E[9], E[10]) + E[11];
      /* pix = cam @ K^T (bmm) ; pix.xy /= pix.z ; pix.xy /= inv_scale */
      const float qx = dot3(cx, cy, cz, K[0], K[1], K[2]);
      const float qy = dot3(cx, cy, cz, K[3], K[4], K[5]);
      const float qz = dot3(cx, cy, cz, K[6], K[7], K[8]);
      const float u = (qx / qz) / isx;
      const float w = (qy / qz) / isy;
      c += (u >= 0.f) && (u <= 1.f) && (w >= 0.f) && (w <= 1.f) && (qz > 0.f);
    }
    count[i] = c;
  }
}

/* raws (K,R,S,4), masks (K,R,S) un-normalised, zs (K,R,S) -> rgb (R,3), depth (R), weights (R,S) */
void oracle_composite_blend(const float* raws, const float* masks, const float* zs, int K, int64_t R, int S,
                            float* rgb, float* depth, float* weights) {
  for (int64_t r = 0; r < R; ++r) {
    float T = 1.f;
    float c[3] = {0.f, 0.f, 0.f};
    float mx = -INFINITY;
    for (int s = 0; s < S; ++s) {
      const int64_t i = r * S + s;
      float msum = 0.f;
      for (int k = 0; k < K; ++k) msum += masks[(int64_t)k * R * S + i];
      float A = 0.f;
      for (int k = 0; k < K; ++k) {
        const float m = masks[(int64_t)k * R * S + i];
        const float wk = msum > 0.f ? m / msum : 1.f / (float)K;
        const float alpha = 1.f - expf(-raws[((int64_t)k * R * S + i) * 4 + 3]);
        A += alpha * wk;
        for (int ch = 0; ch < 3; ++ch) c[ch] += (T * alpha * wk) * raws[((int64_t)k * R * S + i) * 4 + ch];
      }
      weights[i] = A * T;
      if (weights[i] > mx) mx = weights[i];
      T = T * (1.f - A);
    }
    float den = 0.f;
    for (int s = 0; s < S; ++s) { weights[r * S + s] = expf(weights[r * S + s] - mx); den += weights[r * S + s]; }
    float d = 0.f;
    for (int s = 0; s < S; ++s) {
      float zm = 0.f;
      for (int k = 0; k < K; ++k) zm += zs[(int64_t)k * R * S + r * S + s];
      zm /= (float)K;
      weights[r * S + s] /= den;
      d += weights[r * S + s] * zm;
    }
    rgb[3 * r] = c[0]; rgb[3 * r + 1] = c[1]; rgb[3 * r + 2] = c[2];
    depth[r] = d;
  }
}
