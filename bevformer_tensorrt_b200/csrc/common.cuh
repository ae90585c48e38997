// common.cuh — shared device/host helpers for libb200_bev_ops (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "b200_bev_ops.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libb200_bev_ops is written for sm_100a (B200) only"
#endif

namespace b200 {

extern std::atomic<unsigned long long> g_launch_count;

inline int check_launch() {
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return cudaGetLastError() == cudaSuccess ? B200_OK : B200_ERR_LAUNCH;
}

constexpr unsigned kFullMask = 0xffffffffu;

// ---- 128-bit read-only loads -------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg128(const void *p) { return __ldg(reinterpret_cast<const uint4 *>(p)); }
__device__ __forceinline__ uint2 ldg64(const void *p) { return __ldg(reinterpret_cast<const uint2 *>(p)); }
__device__ __forceinline__ uint32_t ldg32(const void *p) { return __ldg(reinterpret_cast<const uint32_t *>(p)); }

// streaming (read-once) variants: do not allocate in L1, evict-first in L2 is left to the default policy
__device__ __forceinline__ uint4 ldg128_stream(const void *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg64_stream(const void *p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg32_stream(const void *p) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void stg128_stream(void *p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

__device__ __forceinline__ float2 h2_to_f2(uint32_t u) {
  return __half22float2(*reinterpret_cast<const __half2 *>(&u));
}
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}

// Blackwell mixed-precision FMA (PTX ISA 8.6, sm_100+): d = a(f16) * b(f16) + c(f32), one rounding, SASS FHFMA.
__device__ __forceinline__ float fma_f32_f16(unsigned short a, unsigned short b, float c) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}

// T2int8<float> of the reference (multiScaleDeformableAttnKernel.cu:51-55): saturate, round half away from zero.
__device__ __forceinline__ int to_int8_sat(float a) {
  a = a > 127.f ? 127.f : a;
  a = a < -128.f ? -128.f : a;
  return static_cast<int>(a + (a > 0.f ? 0.5f : -0.5f));
}

}  // namespace b200
