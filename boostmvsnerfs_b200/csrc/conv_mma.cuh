// Warp-level tensor-core primitives shared by the convolution kernels (conv3d_mma.cu, convT3d_mma.cu).
#pragma once
#include <cuda_fp16.h>

#include <cstdint>

#include "bmv_internal.cuh"

namespace bmv {

// four 8x8 b16 matrices; lane l supplies the row address of matrix l/8, row l%8
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
// D(16x8,f32) += A(16x16,f16,row) * B(16x8,f16,col)
__device__ __forceinline__ void hmma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint2 pack_half4(const float4& v) {
  uint2 pk;
  pk.x = pack_half2_sat(v.x, v.y);
  pk.y = pack_half2_sat(v.z, v.w);
  return pk;
}

}  // namespace bmv
