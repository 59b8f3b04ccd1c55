// The per-sample MLP (reference lib/networks/enerf/nerf.py:29-89) on one 16-sample warp-level MMA tile: shared by
// render_mma.cu (one chain per launch) and render_multi.cu (all K chains in one persistent launch).
// Staged row layout (kStageStride floats): vox 8 | 3 x (f_v: feat 8, rgb 3, dir 4, pad 1).
#pragma once
#include <cuda_fp16.h>

#include "bmv_internal.cuh"

namespace bmv {

constexpr int kStageStride = 72;        // floats per staged sample: vox 8 | 3 x (f_v 15 + pad) ; 72 = 8 mod 32

// fragment-ordered weight blocks (128 words each), in this order
constexpr int BLK_GS = 0;               // global_fc, [var | mean] part : KT=2, NT=4
constexpr int BLK_GV = BLK_GS + 2 * 4;  // global_fc, per-view x part   : KT=1, NT=4
constexpr int BLK_FC = BLK_GV + 1 * 4;  // agg.fc                       : KT=2, NT=2
constexpr int BLK_L0 = BLK_FC + 2 * 2;  // lr0  [pooled | vox]          : KT=2, NT=8
constexpr int BLK_CS = BLK_L0 + 2 * 8;  // color.0 [hid | pooled | vox] : KT=6, NT=8
constexpr int BLK_CV = BLK_CS + 6 * 8;  // color.0 per-view f_v         : KT=1, NT=8
constexpr int NUM_BLK = BLK_CV + 1 * 8;
// fp32 vectors after the blocks (word offsets)
constexpr int V_BG = NUM_BLK * 128;     // global_fc.bias[32]
constexpr int V_WA = V_BG + 32;         // agg_w_fc.weight[32]
constexpr int V_BFC = V_WA + 32;        // fc.bias[16]
constexpr int V_BL = V_BFC + 16;        // lr0.bias[64]
constexpr int V_WS = V_BL + 64;         // sigma.weight[64]
constexpr int V_BC = V_WS + 64;         // color.0.bias[64]
constexpr int V_W2 = V_BC + 64;         // color.2.weight[64]
constexpr int V_WV = V_W2 + 64;         // view_fc.weight[12][4]
constexpr int V_BV = V_WV + 48;         // view_fc.bias[12]
constexpr int V_SC = V_BV + 12;         // agg_w_fc.bias, sigma.bias, color.2.bias, 0
constexpr int MMA_PACK_WORDS = V_SC + 4;

// x = hi + lo, hi = fp16(x), lo = fp16(x - hi) for two values.  The residual x - float(hi) is taken with the sm_100
// mixed-precision FMA (fma.rn.f32.f16: hi * (-1) + x, one FHFMA per value reading the half straight out of the packed
// register) instead of unpack + subtract: 4 instructions per pair instead of 6, bit-identical result.
__device__ __forceinline__ void split_pack(float v0, float v1, uint32_t& hi, uint32_t& lo) {
  float r0, r1;
  asm("{\n\t.reg .b16 l, h, m;\n\t"
      "cvt.rn.f16x2.f32 %0, %4, %3;\n\t"
      "mov.b32 {l, h}, %0;\n\t"
      "mov.b16 m, 0xBC00;\n\t"
      "fma.rn.f32.f16 %1, l, m, %3;\n\t"
      "fma.rn.f32.f16 %2, h, m, %4;\n\t}"
      : "=r"(hi), "=f"(r0), "=f"(r1) : "f"(v0), "f"(v1));
  asm("cvt.rn.f16x2.f32 %0, %2, %1;" : "=r"(lo) : "f"(r0), "f"(r1));
}

struct AFrag { uint32_t hi[4], lo[4]; };

// fr[row 0/1][slot 0..3] = values at (row g / g+8, cols 2t, 2t+1, 2t+8, 2t+9) of a 16-wide K tile
__device__ __forceinline__ AFrag make_afrag(const float (&fr)[2][4]) {
  AFrag a;
  split_pack(fr[0][0], fr[0][1], a.hi[0], a.lo[0]);
  split_pack(fr[1][0], fr[1][1], a.hi[1], a.lo[1]);
  split_pack(fr[0][2], fr[0][3], a.hi[2], a.lo[2]);
  split_pack(fr[1][2], fr[1][3], a.hi[3], a.lo[3]);
  return a;
}

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// c += A * B for one (k-tile, n-tile) with split operands: hi*hi + hi*lo + lo*hi
__device__ __forceinline__ void mma3(float (&c)[4], const AFrag& a, const uint32_t* __restrict__ sW, int blk, int lane) {
  const uint4 b = *reinterpret_cast<const uint4*>(sW + blk * 128 + lane * 4);
  mma16816(c, a.lo, b.x, b.y);
  mma16816(c, a.hi, b.z, b.w);
  mma16816(c, a.hi, b.x, b.y);
}

__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  return v;
}

// Rows g and g+8 of the tile (row0 / row1 = their staged rows).  Result: o[0] = (rgb, sigma) of row g, o[1] of row g+8,
// identical in the 4 lanes of a quad.
__device__ __forceinline__ void mlp_mma_tile(const float* row0, const float* row1, const uint32_t* sW, const float* sV,
                                             int lane, float ba, float bs, float b2, float4 (&o)[2]) {
  constexpr int V = 3;
  const int t = lane & 3;
  const int cols[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
  // per-view feature tiles f_v (cols 0..15 of the view block) and view_fc -> x_v
  AFrag fA[V];
  float x[V][2][4];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const float* r0 = row0 + 8 + v * 16;
    const float* r1 = row1 + 8 + v * 16;
    float fr[2][4];
    const float2 a0 = *reinterpret_cast<const float2*>(r0 + 2 * t), a2 = *reinterpret_cast<const float2*>(r0 + 2 * t + 8);
    const float2 a1 = *reinterpret_cast<const float2*>(r1 + 2 * t), a3 = *reinterpret_cast<const float2*>(r1 + 2 * t + 8);
    fr[0][0] = a0.x; fr[0][1] = a0.y; fr[0][2] = a2.x; fr[0][3] = a2.y;
    fr[1][0] = a1.x; fr[1][1] = a1.y; fr[1][2] = a3.x; fr[1][3] = a3.y;
    fA[v] = make_afrag(fr);
    // x = feat + relu(Wv . dir + bv) for the feature columns (< 11); 0 in the padding columns
    const float d0[4] = {r0[11], r0[12], r0[13], r0[14]};
    const float d1[4] = {r1[11], r1[12], r1[13], r1[14]};
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      const int c = cols[sl];
      if (c < 11) {
        const float4 w = *reinterpret_cast<const float4*>(sV + V_WV + c * 4);
        const float b = sV[V_BV + c];
        const float e0 = fmaf(w.w, d0[3], fmaf(w.z, d0[2], fmaf(w.y, d0[1], fmaf(w.x, d0[0], b))));
        const float e1 = fmaf(w.w, d1[3], fmaf(w.z, d1[2], fmaf(w.y, d1[1], fmaf(w.x, d1[0], b))));
        x[v][0][sl] = fr[0][sl] + fmaxf(e0, 0.f);
        x[v][1][sl] = fr[1][sl] + fmaxf(e1, 0.f);
      } else {
        x[v][0][sl] = 0.f; x[v][1][sl] = 0.f;
      }
    }
  }
  // mean / unbiased variance over the views (elementwise), then global_fc
  float G[V][4][4];                                   // [view][n-tile][c-frag]
  {
    float var[2][4], mean[2][4];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const float m = (x[0][r][sl] + x[1][r][sl] + x[2][r][sl]) * (1.f / 3.f);
        const float e0 = x[0][r][sl] - m, e1 = x[1][r][sl] - m, e2 = x[2][r][sl] - m;
        mean[r][sl] = m;
        var[r][sl] = fmaf(e2, e2, fmaf(e1, e1, e0 * e0)) * 0.5f;
      }
    const AFrag aVar = make_afrag(var), aMean = make_afrag(mean);
    AFrag aX[V];
#pragma unroll
    for (int v = 0; v < V; ++v) aX[v] = make_afrag(x[v]);
    float sh[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float bg0 = sV[V_BG + nt * 8 + 2 * t], bg1 = sV[V_BG + nt * 8 + 2 * t + 1];
      sh[nt][0] = bg0; sh[nt][1] = bg1; sh[nt][2] = bg0; sh[nt][3] = bg1;
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma3(sh[nt], aVar, sW, BLK_GS + 0 * 4 + nt, lane);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) mma3(sh[nt], aMean, sW, BLK_GS + 1 * 4 + nt, lane);
#pragma unroll
    for (int v = 0; v < V; ++v) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) G[v][nt][i] = sh[nt][i];
        mma3(G[v][nt], aX[v], sW, BLK_GV + nt, lane);
      }
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int i = 0; i < 4; ++i) G[v][nt][i] = fmaxf(G[v][nt][i], 0.f);
  }
  // agg_w_fc + softmax over views, im = sum_v w_v G_v
  float im[4][4];
  {
    float lg[V][2];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const float w0 = sV[V_WA + nt * 8 + 2 * t], w1 = sV[V_WA + nt * 8 + 2 * t + 1];
        s0 = fmaf(w1, G[v][nt][1], fmaf(w0, G[v][nt][0], s0));
        s1 = fmaf(w1, G[v][nt][3], fmaf(w0, G[v][nt][2], s1));
      }
      lg[v][0] = fmaxf(quad_sum(s0) + ba, 0.f);
      lg[v][1] = fmaxf(quad_sum(s1) + ba, 0.f);
    }
    float wv[V][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mx = fmaxf(lg[0][r], fmaxf(lg[1][r], lg[2][r]));
      const float e0 = expf(lg[0][r] - mx), e1 = expf(lg[1][r] - mx), e2 = expf(lg[2][r] - mx);
      const float inv = 1.f / (e0 + e1 + e2);
      wv[0][r] = e0 * inv; wv[1][r] = e1 * inv; wv[2][r] = e2 * inv;
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = i >> 1;
        im[nt][i] = fmaf(wv[2][r], G[2][nt][i], fmaf(wv[1][r], G[1][nt][i], wv[0][r] * G[0][nt][i]));
      }
  }
  // fc: 32 -> 16 (+ReLU) ; pooled as the next layer's K tile
  AFrag aPooled;
  {
    AFrag aIm[2];
#pragma unroll
    for (int kt = 0; kt < 2; ++kt) {
      float fr[2][4] = {{im[2 * kt][0], im[2 * kt][1], im[2 * kt + 1][0], im[2 * kt + 1][1]},
                        {im[2 * kt][2], im[2 * kt][3], im[2 * kt + 1][2], im[2 * kt + 1][3]}};
      aIm[kt] = make_afrag(fr);
    }
    float pc[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const float b0 = sV[V_BFC + nt * 8 + 2 * t], b1 = sV[V_BFC + nt * 8 + 2 * t + 1];
      pc[nt][0] = b0; pc[nt][1] = b1; pc[nt][2] = b0; pc[nt][3] = b1;
    }
#pragma unroll
    for (int kt = 0; kt < 2; ++kt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) mma3(pc[nt], aIm[kt], sW, BLK_FC + kt * 2 + nt, lane);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int i = 0; i < 4; ++i) pc[nt][i] = fmaxf(pc[nt][i], 0.f);
    float fr[2][4] = {{pc[0][0], pc[0][1], pc[1][0], pc[1][1]}, {pc[0][2], pc[0][3], pc[1][2], pc[1][3]}};
    aPooled = make_afrag(fr);
  }
  // vox K tile: cols 0..7 = vox, 8..15 = 0
  AFrag aVox;
  {
    const float2 v0 = *reinterpret_cast<const float2*>(row0 + 2 * t), v1 = *reinterpret_cast<const float2*>(row1 + 2 * t);
    float fr[2][4] = {{v0.x, v0.y, 0.f, 0.f}, {v1.x, v1.y, 0.f, 0.f}};
    aVox = make_afrag(fr);
  }
  // lr0: [pooled | vox] -> 64 (+ReLU), sigma = softplus(ws . hid + bs); hid as 4 K tiles
  AFrag aHid[4];
  float sig0 = 0.f, sig1 = 0.f;
  {
    float hc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float b0 = sV[V_BL + nt * 8 + 2 * t], b1 = sV[V_BL + nt * 8 + 2 * t + 1];
      hc[nt][0] = b0; hc[nt][1] = b1; hc[nt][2] = b0; hc[nt][3] = b1;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma3(hc[nt], aPooled, sW, BLK_L0 + 0 * 8 + nt, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) mma3(hc[nt], aVox, sW, BLK_L0 + 1 * 8 + nt, lane);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float w0 = sV[V_WS + nt * 8 + 2 * t], w1 = sV[V_WS + nt * 8 + 2 * t + 1];
#pragma unroll
      for (int i = 0; i < 4; ++i) hc[nt][i] = fmaxf(hc[nt][i], 0.f);
      sig0 = fmaf(w1, hc[nt][1], fmaf(w0, hc[nt][0], sig0));
      sig1 = fmaf(w1, hc[nt][3], fmaf(w0, hc[nt][2], sig1));
    }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      float fr[2][4] = {{hc[2 * kt][0], hc[2 * kt][1], hc[2 * kt + 1][0], hc[2 * kt + 1][1]},
                        {hc[2 * kt][2], hc[2 * kt][3], hc[2 * kt + 1][2], hc[2 * kt + 1][3]}};
      aHid[kt] = make_afrag(fr);
    }
  }
  sig0 = quad_sum(sig0) + bs;
  sig1 = quad_sum(sig1) + bs;
  sig0 = sig0 > 20.f ? sig0 : log1pf(expf(sig0));
  sig1 = sig1 > 20.f ? sig1 : log1pf(expf(sig1));
  // color.0 (shared part once per n-tile, per-view part on top) + color.2 partial dot
  float cl[V][2];
#pragma unroll
  for (int v = 0; v < V; ++v) { cl[v][0] = 0.f; cl[v][1] = 0.f; }
#pragma unroll 1
  for (int n0 = 0; n0 < 8; n0 += 4) {
    float sh[4][4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float b0 = sV[V_BC + (n0 + j) * 8 + 2 * t], b1 = sV[V_BC + (n0 + j) * 8 + 2 * t + 1];
      sh[j][0] = b0; sh[j][1] = b1; sh[j][2] = b0; sh[j][3] = b1;
    }
#pragma unroll
    for (int kt = 0; kt < 4; ++kt)
#pragma unroll
      for (int j = 0; j < 4; ++j) mma3(sh[j], aHid[kt], sW, BLK_CS + kt * 8 + n0 + j, lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) mma3(sh[j], aPooled, sW, BLK_CS + 4 * 8 + n0 + j, lane);
#pragma unroll
    for (int j = 0; j < 4; ++j) mma3(sh[j], aVox, sW, BLK_CS + 5 * 8 + n0 + j, lane);
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float c[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) c[j][i] = sh[j][i];
        mma3(c[j], fA[v], sW, BLK_CV + n0 + j, lane);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float w0 = sV[V_W2 + (n0 + j) * 8 + 2 * t], w1 = sV[V_W2 + (n0 + j) * 8 + 2 * t + 1];
        cl[v][0] = fmaf(w1, fmaxf(c[j][1], 0.f), fmaf(w0, fmaxf(c[j][0], 0.f), cl[v][0]));
        cl[v][1] = fmaf(w1, fmaxf(c[j][3], 0.f), fmaf(w0, fmaxf(c[j][2], 0.f), cl[v][1]));
      }
    }
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    cl[v][0] = fmaxf(quad_sum(cl[v][0]) + b2, 0.f);
    cl[v][1] = fmaxf(quad_sum(cl[v][1]) + b2, 0.f);
  }
  // softmax over views, rgb = sum_v beta_v rgb_v (meaningful in every lane; the caller lets lane t == 0 write rows g, g+8)
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float* rw = r ? row1 : row0;
    const float mx = fmaxf(cl[0][r], fmaxf(cl[1][r], cl[2][r]));
    const float e0 = expf(cl[0][r] - mx), e1 = expf(cl[1][r] - mx), e2 = expf(cl[2][r] - mx);
    const float inv = 1.f / (e0 + e1 + e2);
    o[r].x = (e0 * rw[8 + 0 * 16 + 8] + e1 * rw[8 + 1 * 16 + 8] + e2 * rw[8 + 2 * 16 + 8]) * inv;
    o[r].y = (e0 * rw[8 + 0 * 16 + 9] + e1 * rw[8 + 1 * 16 + 9] + e2 * rw[8 + 2 * 16 + 9]) * inv;
    o[r].z = (e0 * rw[8 + 0 * 16 + 10] + e1 * rw[8 + 1 * 16 + 10] + e2 * rw[8 + 2 * 16 + 10]) * inv;
    o[r].w = r ? sig1 : sig0;
  }
}

}  // namespace bmv
