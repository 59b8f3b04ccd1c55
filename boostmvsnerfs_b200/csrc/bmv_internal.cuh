// Shared host/device helpers for libbmv (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include <nvtx3/nvToolsExt.h>

#include "../../include/bmv.h"

namespace bmv {

void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
    return BMV_ERR_CUDA_LAUNCH;
  }
  return BMV_OK;
}

#define BMV_REQUIRE(cond, code, ...)      \
  do {                                    \
    if (!(cond)) {                        \
      ::bmv::set_error(__VA_ARGS__);      \
      return (code);                      \
    }                                     \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// One NVTX range per C-ABI entry (host side: argument checks + enqueue).  With no profiler attached the NVTX v3
// header-only shim is a load + branch per call; under nsys / ncu --nvtx the frame reads as a list of named libbmv calls.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define BMV_NVTX_RANGE(name) ::bmv::NvtxRange bmv_nvtx_range_(name)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize / carve-out) is a PER-DEVICE attribute: a launcher configures its
// kernel once per device it is used on (a process-wide flag would leave a second GPU of the same process unconfigured).
// Thread-safe: concurrent first calls may both configure (idempotent), none skips it.
struct DeviceOnce {
  std::atomic<uint64_t> mask[2]{};                     // devices 0..127
  // the current device if it has not been configured yet (configure, then call done(dev)); -1 otherwise
  int needed() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 128) return 0;   // let the configure call report it
    return ((mask[dev >> 6].load(std::memory_order_acquire) >> (dev & 63)) & 1u) ? -1 : dev;
  }
  void done(int dev) { mask[dev >> 6].fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ----------------------------------------------------------------------------- device helpers
// Exactly-rounded primitives where the reference has SEPARATE torch ops (no FMA contraction).
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }

// 3-term dot product the way an SGEMM inner loop accumulates it (k ascending, acc starts at 0):
// acc = a0*b0; acc = fma(a1,b1,acc); acc = fma(a2,b2,acc).  Verified bit-identical to torch.bmm
// on CPU (MKL) — tests/test_oracle_golden.py::test_visibility pins the CPU side.
__device__ __forceinline__ float dot3_gemm(float a0, float a1, float a2, float b0, float b1, float b2) {
  float acc = __fmul_rn(a0, b0);
  acc = __fmaf_rn(a1, b1, acc);
  acc = __fmaf_rn(a2, b2, acc);
  return acc;
}
__device__ __forceinline__ float dot4_gemm(float a0, float a1, float a2, float a3,
                                           float b0, float b1, float b2, float b3) {
  float acc = __fmul_rn(a0, b0);
  acc = __fmaf_rn(a1, b1, acc);
  acc = __fmaf_rn(a2, b2, acc);
  acc = __fmaf_rn(a3, b3, acc);
  return acc;
}

// fp32 -> fp16 conversions of every STORED activation / volume element saturate at +-65504 instead of producing inf
// (one F2FP.SATFINITE instruction; NaN stays NaN): an out-of-range value then costs accuracy, not a NaN-poisoned frame.
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ __half half_sat(float v) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}

// ATen grid_sampler unnormalize with align_corners=True: ((g + 1) / 2) * (size - 1).
// (x / 2 is written x * 0.5: the same correctly rounded value for every x, without the IEEE division sequence.)
__device__ __forceinline__ float unnormalize_ac(float g, int size) {
  return mul_rn(mul_rn(add_rn(g, 1.f), 0.5f), (float)(size - 1));
}

// A coordinate ATen's CUDA sampler would send to "-100" (non-finite or outside int range).
__device__ __forceinline__ bool coord_ok(float v) { return fabsf(v) < 1.0e9f; }  // false for NaN/inf

// align_corners=True bilinear upsample source position (ATen upsample_bilinear2d):
// scale = (in-1)/(out-1) in fp32, src = scale*dst, i0 = (int)src, i1 = i0 + (i0 < in-1), lambda = src-i0.
struct UpCoord { int i0, i1; float l0, l1; };
__device__ __forceinline__ float up_scale(int in_size, int out_size) {
  return out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
}
__device__ __forceinline__ UpCoord up_coord_scaled(int dst, int in_size, float scale);
__device__ __forceinline__ UpCoord up_coord(int dst, int in_size, int out_size) {
  return up_coord_scaled(dst, in_size, up_scale(in_size, out_size));
}
// same with the (loop-invariant) scale hoisted by the caller
__device__ __forceinline__ UpCoord up_coord_scaled(int dst, int in_size, float scale) {
  float src = mul_rn(scale, (float)dst);
  UpCoord u;
  u.i0 = (int)src;
  u.i1 = u.i0 + ((u.i0 < in_size - 1) ? 1 : 0);
  u.l1 = sub_rn(src, (float)u.i0);
  u.l0 = sub_rn(1.f, u.l1);
  return u;
}
__device__ __forceinline__ float up_sample(const float* __restrict__ m, int w_in, const UpCoord& uy, const UpCoord& ux) {
  float a = __ldg(m + (int64_t)uy.i0 * w_in + ux.i0), b = __ldg(m + (int64_t)uy.i0 * w_in + ux.i1);
  float c = __ldg(m + (int64_t)uy.i1 * w_in + ux.i0), d = __ldg(m + (int64_t)uy.i1 * w_in + ux.i1);
  // ATen: h0l*(w0l*a + w1l*b) + h1l*(w0l*c + w1l*d)
  float top = add_rn(mul_rn(ux.l0, a), mul_rn(ux.l1, b));
  float bot = add_rn(mul_rn(ux.l0, c), mul_rn(ux.l1, d));
  return add_rn(mul_rn(uy.l0, top), mul_rn(uy.l1, bot));
}

}  // namespace bmv
