// K4: alpha compositing, single volume and visibility-weighted K-volume blend.
// Reference semantics: raw2outputs (lib/networks/enerf/utils.py:605-637), merge_mlp_outputs
// (lib/networks/boost_enerf/network.py:163-170), raw2outputs_blend (lib/networks/enerf/utils.py:639-667).
#include "bmv_internal.cuh"

namespace bmv {

constexpr int kMaxSerialS = 16;   // rays with <= 16 samples: one thread per ray, registers only

// ---------------------------------------------------------------- K-blend, short rays (ENeRF: S=2..8)
// One thread per ray, nothing indexed by k is kept: per sample the K masks are summed first, then
// one pass over k accumulates A = sum_k alpha_k w_k, the colour and mean z.  Inputs are read once
// from HBM (the second mask read hits L1): K*S*(16+4+4) B in, 12+4+4S B out per ray.
template <int MAXS>
__global__ void __launch_bounds__(256) composite_blend_serial_kernel(bmv_composite_blend_params p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  const int K = p.K, S = p.S;
  const float invK = div_rn(1.f, (float)K);
  float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
  float wts[MAXS], zm[MAXS];
#pragma unroll
  for (int s = 0; s < MAXS; ++s) {
    if (s >= S) break;
    const int64_t i = r * S + s;
    float msum = 0.f;
    for (int k = 0; k < K; ++k) msum = add_rn(msum, __ldg(p.mask[k] + i));
    float A = 0.f, zacc = 0.f, sr = 0.f, sg = 0.f, sb = 0.f;
    for (int k = 0; k < K; ++k) {
      const float4 raw = __ldg(reinterpret_cast<const float4*>(p.raw[k]) + i);
      const float wk = msum > 0.f ? div_rn(__ldg(p.mask[k] + i), msum) : invK;
      const float aw = mul_rn(sub_rn(1.f, expf(-raw.w)), wk);
      A = add_rn(A, aw);
      sr = fmaf(aw, raw.x, sr); sg = fmaf(aw, raw.y, sg); sb = fmaf(aw, raw.z, sb);
      zacc = add_rn(zacc, __ldg(p.z[k] + i));
    }
    cr = fmaf(T, sr, cr); cg = fmaf(T, sg, cg); cb = fmaf(T, sb, cb);
    wts[s] = mul_rn(A, T);
    zm[s] = div_rn(zacc, (float)K);
    T = mul_rn(T, sub_rn(1.f, A));       // cumprod([1, 1-A]) — no epsilon in the blend
  }
  // weights <- softmax_s(A*T); depth = sum softmax * mean_k z
  float mx = -INFINITY;
#pragma unroll
  for (int s = 0; s < MAXS; ++s) { if (s >= S) break; mx = fmaxf(mx, wts[s]); }
  float den = 0.f;
#pragma unroll
  for (int s = 0; s < MAXS; ++s) { if (s >= S) break; wts[s] = expf(wts[s] - mx); den += wts[s]; }
  float depth = 0.f;
#pragma unroll
  for (int s = 0; s < MAXS; ++s) {
    if (s >= S) break;
    const float w = div_rn(wts[s], den);
    if (p.weights) p.weights[r * S + s] = w;
    depth = add_rn(depth, mul_rn(w, zm[s]));
  }
  if (p.rgb) { p.rgb[r * 3] = cr; p.rgb[r * 3 + 1] = cg; p.rgb[r * 3 + 2] = cb; }
  if (p.depth) p.depth[r] = depth;
}

// S == 2 (the ENeRF evaluation configs): both samples of a ray are fetched with one 8-byte load per
// (k, mask|z) and two 16-byte loads per (k, raw) — 4K vector loads in flight per thread instead of 8K
// scalar ones; same arithmetic as the generic kernel.
__global__ void __launch_bounds__(128) composite_blend_s2_kernel(bmv_composite_blend_params p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  const int K = p.K;
  const float invK = div_rn(1.f, (float)K);
  float m0 = 0.f, m1 = 0.f;
  for (int k = 0; k < K; ++k) {
    const float2 m = __ldg(reinterpret_cast<const float2*>(p.mask[k]) + r);
    m0 = add_rn(m0, m.x); m1 = add_rn(m1, m.y);
  }
  float A0 = 0.f, A1 = 0.f, z0 = 0.f, z1 = 0.f;
  float r0 = 0.f, g0 = 0.f, b0 = 0.f, r1 = 0.f, g1 = 0.f, b1 = 0.f;
  for (int k = 0; k < K; ++k) {
    const float4 ra = __ldg(reinterpret_cast<const float4*>(p.raw[k]) + 2 * r);
    const float4 rb = __ldg(reinterpret_cast<const float4*>(p.raw[k]) + 2 * r + 1);
    const float2 m = __ldg(reinterpret_cast<const float2*>(p.mask[k]) + r);
    const float2 z = __ldg(reinterpret_cast<const float2*>(p.z[k]) + r);
    const float w0 = m0 > 0.f ? div_rn(m.x, m0) : invK, w1 = m1 > 0.f ? div_rn(m.y, m1) : invK;
    const float a0 = mul_rn(sub_rn(1.f, expf(-ra.w)), w0), a1 = mul_rn(sub_rn(1.f, expf(-rb.w)), w1);
    A0 = add_rn(A0, a0); A1 = add_rn(A1, a1);
    r0 = fmaf(a0, ra.x, r0); g0 = fmaf(a0, ra.y, g0); b0 = fmaf(a0, ra.z, b0);
    r1 = fmaf(a1, rb.x, r1); g1 = fmaf(a1, rb.y, g1); b1 = fmaf(a1, rb.z, b1);
    z0 = add_rn(z0, z.x); z1 = add_rn(z1, z.y);
  }
  const float T1 = sub_rn(1.f, A0);                    // T0 = 1
  const float wt0 = A0, wt1 = mul_rn(A1, T1);
  const float mx = fmaxf(wt0, wt1);
  const float e0 = expf(wt0 - mx), e1 = expf(wt1 - mx);
  const float den = e0 + e1;
  const float s0 = div_rn(e0, den), s1 = div_rn(e1, den);
  if (p.weights) reinterpret_cast<float2*>(p.weights)[r] = make_float2(s0, s1);
  if (p.depth) p.depth[r] = add_rn(mul_rn(s0, div_rn(z0, (float)K)), mul_rn(s1, div_rn(z1, (float)K)));
  if (p.rgb) {
    p.rgb[r * 3] = fmaf(T1, r1, r0); p.rgb[r * 3 + 1] = fmaf(T1, g1, g0); p.rgb[r * 3 + 2] = fmaf(T1, b1, b0);
  }
}

// ---------------------------------------------------------------- warp-per-ray variants (long rays)
// Lanes stride over samples; transmittance is an exclusive multiplicative warp scan carried across
// 32-sample segments.  Used for MVSNeRF-style rays (S = 32..128+).
__device__ __forceinline__ float warp_excl_prod(float v, int lane, float& total) {
  float inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc *= n;
  }
  total = __shfl_sync(0xffffffffu, inc, 31);
  float ex = __shfl_up_sync(0xffffffffu, inc, 1);
  return lane == 0 ? 1.f : ex;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__global__ void __launch_bounds__(256) composite_blend_warp_kernel(bmv_composite_blend_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= p.R) return;
  const int K = p.K, S = p.S;
  float carry = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
  float mx = -INFINITY;
  // pass 1: transmittance scan, colour, un-normalised weights (stashed in p.weights)
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool on = s < S;
    float A = 0.f, cw_r = 0.f, cw_g = 0.f, cw_b = 0.f;
    if (on) {
      float msum = 0.f;
      for (int k = 0; k < K; ++k) msum = add_rn(msum, __ldg(p.mask[k] + r * S + s));
      for (int k = 0; k < K; ++k) {
        const float4 raw = __ldg(reinterpret_cast<const float4*>(p.raw[k]) + r * S + s);
        const float mk = __ldg(p.mask[k] + r * S + s);
        const float wk = msum > 0.f ? div_rn(mk, msum) : div_rn(1.f, (float)K);
        const float a = sub_rn(1.f, expf(-raw.w));
        const float aw = mul_rn(a, wk);
        A = add_rn(A, aw);
        cw_r = fmaf(aw, raw.x, cw_r); cw_g = fmaf(aw, raw.y, cw_g); cw_b = fmaf(aw, raw.z, cw_b);
      }
    }
    float total;
    const float T = carry * warp_excl_prod(on ? 1.f - A : 1.f, lane, total);
    carry *= total;
    if (on) {
      cr = fmaf(T, cw_r, cr); cg = fmaf(T, cw_g, cg); cb = fmaf(T, cw_b, cb);
      const float w = A * T;
      mx = fmaxf(mx, w);
      p.weights[r * S + s] = w;
    }
  }
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
  mx = warp_max(mx);
  // pass 2: softmax over samples + depth (re-reads this warp's own stores: L1/L2 hits)
  float den = 0.f;
  for (int s = lane; s < S; s += 32) den += expf(p.weights[r * S + s] - mx);
  den = warp_sum(den);
  float depth = 0.f;
  for (int s = lane; s < S; s += 32) {
    const float w = div_rn(expf(p.weights[r * S + s] - mx), den);
    p.weights[r * S + s] = w;
    float zacc = 0.f;
    for (int k = 0; k < K; ++k) zacc += __ldg(p.z[k] + r * S + s);
    depth = fmaf(w, zacc / (float)K, depth);
  }
  depth = warp_sum(depth);
  if (lane == 0) {
    if (p.rgb) { p.rgb[r * 3] = cr; p.rgb[r * 3 + 1] = cg; p.rgb[r * 3 + 2] = cb; }
    if (p.depth) p.depth[r] = depth;
  }
}

// ---------------------------------------------------------------- single-volume compositing
__global__ void __launch_bounds__(256) composite_serial_kernel(bmv_composite_params p) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= p.R) return;
  const int S = p.S;
  float T = 1.f, cr = 0.f, cg = 0.f, cb = 0.f;
  float wts[kMaxSerialS];
#pragma unroll
  for (int s = 0; s < kMaxSerialS; ++s) {
    if (s >= S) break;
    const float4 raw = __ldg(reinterpret_cast<const float4*>(p.raw) + r * S + s);
    const float a = sub_rn(1.f, expf(-raw.w));
    const float w = mul_rn(a, T);
    cr = add_rn(cr, mul_rn(w, raw.x)); cg = add_rn(cg, mul_rn(w, raw.y)); cb = add_rn(cb, mul_rn(w, raw.z));
    wts[s] = w;
    T = mul_rn(T, add_rn(sub_rn(1.f, a), 1e-10f));   // cumprod(1 - alpha + 1e-10)
  }
  float depth = 0.f;
  if (p.z) {
    float mx = -INFINITY, den = 0.f;
#pragma unroll
    for (int s = 0; s < kMaxSerialS; ++s) { if (s >= S) break; mx = fmaxf(mx, wts[s]); }
#pragma unroll
    for (int s = 0; s < kMaxSerialS; ++s) { if (s >= S) break; wts[s] = expf(wts[s] - mx); den += wts[s]; }
#pragma unroll
    for (int s = 0; s < kMaxSerialS; ++s) {
      if (s >= S) break;
      wts[s] = div_rn(wts[s], den);
      depth = add_rn(depth, mul_rn(wts[s], __ldg(p.z + r * S + s)));
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int s = 0; s < kMaxSerialS; ++s) {
    if (s >= S) break;
    if (p.weights) p.weights[r * S + s] = wts[s];
    acc = add_rn(acc, wts[s]);
  }
  if (p.white_bkgd) { const float bg = sub_rn(1.f, acc); cr = add_rn(cr, bg); cg = add_rn(cg, bg); cb = add_rn(cb, bg); }
  if (p.rgb) { p.rgb[r * 3] = cr; p.rgb[r * 3 + 1] = cg; p.rgb[r * 3 + 2] = cb; }
  if (p.depth && p.z) p.depth[r] = depth;
}

__global__ void __launch_bounds__(256) composite_warp_kernel(bmv_composite_params p) {
  const int lane = threadIdx.x & 31;
  const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= p.R) return;
  const int S = p.S;
  float carry = 1.f, cr = 0.f, cg = 0.f, cb = 0.f, mx = -INFINITY;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    const bool on = s < S;
    float4 raw = make_float4(0.f, 0.f, 0.f, 0.f);
    float a = 0.f;
    if (on) { raw = __ldg(reinterpret_cast<const float4*>(p.raw) + r * S + s); a = sub_rn(1.f, expf(-raw.w)); }
    float total;
    const float T = carry * warp_excl_prod(on ? add_rn(sub_rn(1.f, a), 1e-10f) : 1.f, lane, total);
    carry *= total;
    if (on) {
      const float w = a * T;
      cr = fmaf(w, raw.x, cr); cg = fmaf(w, raw.y, cg); cb = fmaf(w, raw.z, cb);
      mx = fmaxf(mx, w);
      p.weights[r * S + s] = w;
    }
  }
  cr = warp_sum(cr); cg = warp_sum(cg); cb = warp_sum(cb);
  float depth = 0.f, acc = 0.f;
  if (p.z) {
    mx = warp_max(mx);
    float den = 0.f;
    for (int s = lane; s < S; s += 32) den += expf(p.weights[r * S + s] - mx);
    den = warp_sum(den);
    for (int s = lane; s < S; s += 32) {
      const float w = div_rn(expf(p.weights[r * S + s] - mx), den);
      p.weights[r * S + s] = w;
      acc += w;
      depth = fmaf(w, __ldg(p.z + r * S + s), depth);
    }
    depth = warp_sum(depth);
  } else {
    for (int s = lane; s < S; s += 32) acc += p.weights[r * S + s];
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    if (p.white_bkgd) { const float bg = 1.f - acc; cr += bg; cg += bg; cb += bg; }
    if (p.rgb) { p.rgb[r * 3] = cr; p.rgb[r * 3 + 1] = cg; p.rgb[r * 3 + 2] = cb; }
    if (p.depth && p.z) p.depth[r] = depth;
  }
}

}  // namespace bmv

extern "C" BMV_API int bmv_composite_blend(const bmv_composite_blend_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_composite_blend");
  using namespace bmv;
  BMV_REQUIRE(p != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_composite_blend: null params");
  BMV_REQUIRE(p->K >= 1 && p->K <= BMV_MAX_VOLUMES, BMV_ERR_INVALID_ARGUMENT, "bmv_composite_blend: K=%d out of 1..%d",
              p->K, BMV_MAX_VOLUMES);
  BMV_REQUIRE(p->S >= 1 && p->R >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_composite_blend: bad S/R");
  if (p->R == 0) return BMV_OK;                       // empty ray set: nothing to enqueue
  for (int k = 0; k < p->K; ++k)
    BMV_REQUIRE(p->raw[k] && p->mask[k] && p->z[k], BMV_ERR_INVALID_ARGUMENT, "bmv_composite_blend: null input %d", k);
  cudaStream_t st = (cudaStream_t)stream;
  bool aligned8 = ((uintptr_t)p->weights & 7) == 0;
  for (int k = 0; k < p->K; ++k)
    aligned8 = aligned8 && ((uintptr_t)p->mask[k] & 7) == 0 && ((uintptr_t)p->z[k] & 7) == 0 && ((uintptr_t)p->raw[k] & 15) == 0;
  if (p->S == 2 && aligned8) {
    composite_blend_s2_kernel<<<(unsigned)ceil_div64(p->R, 128), 128, 0, st>>>(*p);
  } else if (p->S <= 2) {
    composite_blend_serial_kernel<2><<<(unsigned)ceil_div64(p->R, 256), 256, 0, st>>>(*p);
  } else if (p->S <= 8) {
    composite_blend_serial_kernel<8><<<(unsigned)ceil_div64(p->R, 256), 256, 0, st>>>(*p);
  } else if (p->S <= kMaxSerialS) {
    composite_blend_serial_kernel<kMaxSerialS><<<(unsigned)ceil_div64(p->R, 256), 256, 0, st>>>(*p);
  } else {
    BMV_REQUIRE(p->weights != nullptr, BMV_ERR_INVALID_ARGUMENT,
                "bmv_composite_blend: weights output is required when S > %d (used as scratch)", kMaxSerialS);
    composite_blend_warp_kernel<<<(unsigned)ceil_div64(p->R * 32, 256), 256, 0, st>>>(*p);
  }
  return check_launch("bmv_composite_blend");
}

extern "C" BMV_API int bmv_composite(const bmv_composite_params* p, bmv_stream_t stream) {
  BMV_NVTX_RANGE("bmv_composite");
  using namespace bmv;
  BMV_REQUIRE(p != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_composite: null params");
  BMV_REQUIRE(p->S >= 1 && p->R >= 0, BMV_ERR_INVALID_ARGUMENT, "bmv_composite: bad S/R");
  if (p->R == 0) return BMV_OK;                       // empty ray set: nothing to enqueue
  BMV_REQUIRE(p->raw != nullptr, BMV_ERR_INVALID_ARGUMENT, "bmv_composite: null input");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->S <= kMaxSerialS) {
    composite_serial_kernel<<<(unsigned)ceil_div64(p->R, 256), 256, 0, st>>>(*p);
  } else {
    BMV_REQUIRE(p->weights != nullptr, BMV_ERR_INVALID_ARGUMENT,
                "bmv_composite: weights output is required when S > %d (used as scratch)", kMaxSerialS);
    composite_warp_kernel<<<(unsigned)ceil_div64(p->R * 32, 256), 256, 0, st>>>(*p);
  }
  return check_launch("bmv_composite");
}
