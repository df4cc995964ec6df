// Shared helpers for libkvc_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define KVC_BUILDING_LIB 1
#include "kvc_b200.h"

namespace kvc {

constexpr int kWarp = 32;

// ---- error plumbing (thread-local message, negative status codes) ----------
void set_error(const char* fmt, ...);

#define KVC_FAIL(code, ...)        \
  do {                             \
    ::kvc::set_error(__VA_ARGS__); \
    return (code);                 \
  } while (0)

#define KVC_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) KVC_FAIL(code, __VA_ARGS__); \
  } while (0)

// Launch-error check only: never synchronises (the reference ops are async).
#define KVC_CHECK_LAUNCH(what)                                              \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess)                                                 \
      KVC_FAIL(KVC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__));      \
  } while (0)

inline cudaStream_t as_stream(kvc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// SM count of the current device (148 on a B200), queried once per device; grids are sized in multiples of it
inline int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- streaming loads / stores: this path has no reuse, keep L1 out of it ----
__device__ __forceinline__ uint4 ld_stream16(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ld_stream8(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
               : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace kvc
