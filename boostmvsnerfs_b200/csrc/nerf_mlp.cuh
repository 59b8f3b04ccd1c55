// Per-sample ENeRF MLP ("NeRF" + "Agg", reference lib/networks/enerf/nerf.py:6-89) evaluated
// entirely in registers, weights broadcast from shared memory as float4.
//
// the reference materialises (B,P,V,103)-shaped activations in HBM (1.4 GB per
// chain at 960x544) and runs ~25 cuBLAS/elementwise launches; measured on B200 that is 61 % of the
// frame (profiles/round1_stage_times.md).  Here one thread owns one sample: 45+8 inputs in,
// 4 outputs out, 14.6 k FMA in between.  Arithmetic is fp32 FMA (no TF32), so results agree with
// the cuBLAS fp32 path to ~1e-6.
//
// Packed weight layout (floats; produced by boostmvsnerfs_b200/mlp_pack.py, F = feat_ch = Cf+3, V views):
//   [VIEW ]  F rows x 8      : view_fc.weight[j][0:4], view_fc.bias[j], 0,0,0
//   [GLOB ]  32 rows x GROW  : global_fc.weight[j][0:F] (pad to FP), [F:3F] (pad to 2*FP), bias[j], agg_w_fc.weight[j], 0,0
//   [AGGB ]  4               : agg_w_fc.bias, 0,0,0
//   [FC   ]  32 rows x 16    : fc.weight^T (input-major), then 16 : fc.bias
//   [LR0  ]  64 rows x 28    : lr0.weight[j][0:24], lr0.bias[j], sigma.weight[j], 0,0
//   [SIGB ]  4               : sigma.bias, 0,0,0
//   [COL  ]  64 rows x CROW  : color.0.weight[j][0:88] , [88:88+F+4] (pad to FVP), color.0.bias[j], color.2.weight[j], 0,0
//   [COLB ]  4               : color.2.bias, 0,0,0
#pragma once
#include "bmv_internal.cuh"

namespace bmv {

template <int F>
struct MlpLayout {
  static constexpr int FP = (F + 3) / 4 * 4;            // padded per-view feature width
  static constexpr int FV = F + 4;                      // per-view row: feat + dir4
  static constexpr int FVP = (FV + 3) / 4 * 4;
  static constexpr int GROW = FP + 2 * FP + 4;          // x | var,mean | bias, agg_w, 0, 0
  static constexpr int CROW = 88 + FVP + 4;
  static constexpr int OFF_VIEW = 0;
  static constexpr int OFF_GLOB = OFF_VIEW + F * 8;
  static constexpr int OFF_AGGB = OFF_GLOB + 32 * GROW;
  static constexpr int OFF_FC = OFF_AGGB + 4;
  static constexpr int OFF_FCB = OFF_FC + 32 * 16;
  static constexpr int OFF_LR0 = OFF_FCB + 16;
  static constexpr int OFF_SIGB = OFF_LR0 + 64 * 28;
  static constexpr int OFF_COL = OFF_SIGB + 4;
  static constexpr int OFF_COLB = OFF_COL + 64 * CROW;
  static constexpr int TOTAL = OFF_COLB + 4;            // floats, multiple of 4
};

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// acc[0..3] += w . x[0..3] lane-wise: four INDEPENDENT FMA chains per dot product, so a warp keeps
// 4 FMAs in flight per weight quad instead of one 4-deep dependent chain (the v1 kernel stalled on
// exactly that: short-scoreboard/FMA-latency bound at 3 warps per scheduler, profiles/round1b).
__device__ __forceinline__ void fma4(float (&acc)[4], const float4 w, float x0, float x1, float x2, float x3) {
  acc[0] = fmaf(w.x, x0, acc[0]); acc[1] = fmaf(w.y, x1, acc[1]);
  acc[2] = fmaf(w.z, x2, acc[2]); acc[3] = fmaf(w.w, x3, acc[3]);
}
__device__ __forceinline__ float sum4(const float (&a)[4]) { return (a[0] + a[1]) + (a[2] + a[3]); }

// f[v][0:F] = fetched per-view features (image feature channels + rgb), f[v][F:F+4] = direction
// features; vox[8] = cost-volume feature.  Returns (r,g,b,sigma).
template <int F, int V>
__device__ __forceinline__ float4 nerf_mlp_eval(const float* __restrict__ sw, const float (&vox)[8],
                                                const float (&f)[V][F + 4]) {
  using L = MlpLayout<F>;
  static_assert(V >= 2 && V <= 4, "views per volume");
  // ---- Agg.view_fc: x_v = feat_v + relu(Wv dir_v + bv)                      (nerf.py:75-77)
  float x[V][L::FP];
#pragma unroll
  for (int j = 0; j < L::FP; ++j) {
    if (j < F) {
      const float4 w = lds4(sw + L::OFF_VIEW + j * 8);
      const float b = sw[L::OFF_VIEW + j * 8 + 4];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        float a = fmaf(w.x, f[v][F], b);
        a = fmaf(w.y, f[v][F + 1], a);
        a = fmaf(w.z, f[v][F + 2], a);
        a = fmaf(w.w, f[v][F + 3], a);
        x[v][j] = f[v][j] + fmaxf(a, 0.f);
      }
    } else {
#pragma unroll
      for (int v = 0; v < V; ++v) x[v][j] = 0.f;
    }
  }
  // ---- unbiased variance and mean over the views                             (nerf.py:81-82)
  float vm[2 * L::FP];
#pragma unroll
  for (int j = 0; j < L::FP; ++j) {
    float m = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) m += x[v][j];
    m *= (1.f / V);
    float s = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) { const float d = x[v][j] - m; s = fmaf(d, d, s); }
    vm[j] = s * (1.f / (V - 1));
    vm[L::FP + j] = m;
  }
  // ---- global_fc + agg_w_fc, pass 1: softmax logits over views                (nerf.py:84-86)
  float logit[V];
  const float aggb = sw[L::OFF_AGGB];
#pragma unroll
  for (int v = 0; v < V; ++v) logit[v] = aggb;
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    const float* row = sw + L::OFF_GLOB + j * L::GROW;
    const float4 tail = lds4(row + 3 * L::FP);           // bias, agg_w
    float sh[4] = {tail.x, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2 * L::FP; i += 4) fma4(sh, lds4(row + L::FP + i), vm[i], vm[i + 1], vm[i + 2], vm[i + 3]);
    float ga[V][4];
#pragma unroll
    for (int v = 0; v < V; ++v) { ga[v][0] = 0.f; ga[v][1] = 0.f; ga[v][2] = 0.f; ga[v][3] = 0.f; }
#pragma unroll
    for (int i = 0; i < L::FP; i += 4) {
      const float4 w = lds4(row + i);
#pragma unroll
      for (int v = 0; v < V; ++v) fma4(ga[v], w, x[v][i], x[v][i + 1], x[v][i + 2], x[v][i + 3]);
    }
    const float shared = sum4(sh);
#pragma unroll
    for (int v = 0; v < V; ++v) logit[v] = fmaf(tail.y, fmaxf(shared + sum4(ga[v]), 0.f), logit[v]);
  }
  float wsm[V];
  {
    float mx = 0.f;                                      // logits are post-ReLU (>= 0)
#pragma unroll
    for (int v = 0; v < V; ++v) { logit[v] = fmaxf(logit[v], 0.f); mx = fmaxf(mx, logit[v]); }
    float den = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) { wsm[v] = expf(logit[v] - mx); den += wsm[v]; }
    const float inv = 1.f / den;
#pragma unroll
    for (int v = 0; v < V; ++v) wsm[v] *= inv;
  }
  // ---- pass 2: recompute global_fc row j, pool over views, accumulate fc      (nerf.py:87-88)
  float base[24];
#pragma unroll
  for (int c = 0; c < 8; ++c) base[c] = vox[c];
#pragma unroll
  for (int o = 0; o < 16; ++o) base[8 + o] = sw[L::OFF_FCB + o];
#pragma unroll 1
  for (int j = 0; j < 32; ++j) {
    const float* row = sw + L::OFF_GLOB + j * L::GROW;
    float sh[4] = {row[3 * L::FP], 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 2 * L::FP; i += 4) fma4(sh, lds4(row + L::FP + i), vm[i], vm[i + 1], vm[i + 2], vm[i + 3]);
    float ga[V][4];
#pragma unroll
    for (int v = 0; v < V; ++v) { ga[v][0] = 0.f; ga[v][1] = 0.f; ga[v][2] = 0.f; ga[v][3] = 0.f; }
#pragma unroll
    for (int i = 0; i < L::FP; i += 4) {
      const float4 w = lds4(row + i);
#pragma unroll
      for (int v = 0; v < V; ++v) fma4(ga[v], w, x[v][i], x[v][i + 1], x[v][i + 2], x[v][i + 3]);
    }
    const float shared = sum4(sh);
    float im = 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) im = fmaf(wsm[v], fmaxf(shared + sum4(ga[v]), 0.f), im);
    const float* fc = sw + L::OFF_FC + j * 16;
#pragma unroll
    for (int o = 0; o < 16; o += 4) {
      const float4 w = lds4(fc + o);
      base[8 + o] = fmaf(w.x, im, base[8 + o]); base[9 + o] = fmaf(w.y, im, base[9 + o]);
      base[10 + o] = fmaf(w.z, im, base[10 + o]); base[11 + o] = fmaf(w.w, im, base[11 + o]);
    }
  }
#pragma unroll
  for (int o = 0; o < 16; ++o) base[8 + o] = fmaxf(base[8 + o], 0.f);
  // ---- lr0 (24 -> 64, ReLU) and sigma (64 -> 1, softplus)                     (nerf.py:33-37)
  float hid[64];
  float sig = sw[L::OFF_SIGB];
#pragma unroll
  for (int j = 0; j < 64; ++j) {
    const float* row = sw + L::OFF_LR0 + j * 28;
    const float4 tail = lds4(row + 24);                  // bias, sigma weight
    float a4[4] = {tail.x, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 24; i += 4) fma4(a4, lds4(row + i), base[i], base[i + 1], base[i + 2], base[i + 3]);
    hid[j] = fmaxf(sum4(a4), 0.f);
    sig = fmaf(tail.y, hid[j], sig);
  }
  sig = sig > 20.f ? sig : log1pf(expf(sig));            // nn.Softplus(beta=1, threshold=20)
  // ---- color: Linear(88+F+4 -> 64) + ReLU + Linear(64 -> 1) + ReLU, softmax over views (nerf.py:38-42)
  float cl[V];
  const float colb = sw[L::OFF_COLB];
#pragma unroll
  for (int v = 0; v < V; ++v) cl[v] = colb;
#pragma unroll 2
  for (int j = 0; j < 64; ++j) {
    const float* row = sw + L::OFF_COL + j * L::CROW;
    const float4 tail = lds4(row + 88 + L::FVP);         // bias, color.2 weight
    float sh[4] = {tail.x, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 64; i += 4) fma4(sh, lds4(row + i), hid[i], hid[i + 1], hid[i + 2], hid[i + 3]);
#pragma unroll
    for (int i = 0; i < 24; i += 4) fma4(sh, lds4(row + 64 + i), base[i], base[i + 1], base[i + 2], base[i + 3]);
    float av[V][4];
#pragma unroll
    for (int v = 0; v < V; ++v) { av[v][0] = 0.f; av[v][1] = 0.f; av[v][2] = 0.f; av[v][3] = 0.f; }
#pragma unroll
    for (int i = 0; i < L::FVP; i += 4) {
      const float4 w = lds4(row + 88 + i);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        av[v][0] = fmaf(w.x, f[v][i], av[v][0]);
        if (i + 1 < L::FV) av[v][1] = fmaf(w.y, f[v][i + 1], av[v][1]);
        if (i + 2 < L::FV) av[v][2] = fmaf(w.z, f[v][i + 2], av[v][2]);
        if (i + 3 < L::FV) av[v][3] = fmaf(w.w, f[v][i + 3], av[v][3]);
      }
    }
    const float shared = sum4(sh);
#pragma unroll
    for (int v = 0; v < V; ++v) cl[v] = fmaf(tail.y, fmaxf(shared + sum4(av[v]), 0.f), cl[v]);
  }
  float r = 0.f, g = 0.f, b = 0.f;
  {
    float mx = 0.f, den = 0.f, e[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { cl[v] = fmaxf(cl[v], 0.f); mx = fmaxf(mx, cl[v]); }
#pragma unroll
    for (int v = 0; v < V; ++v) { e[v] = expf(cl[v] - mx); den += e[v]; }
    const float inv = 1.f / den;
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const float wv = e[v] * inv;
      r = fmaf(wv, f[v][F - 3], r); g = fmaf(wv, f[v][F - 2], g); b = fmaf(wv, f[v][F - 1], b);
    }
  }
  return make_float4(r, g, b, sig);
}

}  // namespace bmv
