// Shared device helpers for the clsr_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define CLSR_DEVINL __device__ __forceinline__

namespace clsr {

constexpr int kWarp = 32;

CLSR_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
CLSR_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
CLSR_DEVINL float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// Accurate variants (the reference computes these in fp32 with Eigen's full-precision kernels).
CLSR_DEVINL float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

// 16-byte streaming load that does not allocate in L1 (rows are touched once per kernel).
CLSR_DEVINL float4 ldg_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
CLSR_DEVINL void stg_stream(float4* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
// Vector reduction to global memory (sm_90+): one 16-byte RED instead of four.
CLSR_DEVINL void red_add_v4(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
CLSR_DEVINL void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

}  // namespace clsr
