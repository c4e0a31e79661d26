// dy_common.cuh — error plumbing and small device helpers shared by all kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "../../include/droneyolo.h"

namespace dy {

// Thread-local error text returned by dy_last_error().
char* err_buf();
int fail(int code, const char* fmt, ...);

#define DY_CHECK_ARG(cond, ...)                                    \
  do {                                                             \
    if (!(cond)) return ::dy::fail(DY_ERR_INVALID, __VA_ARGS__);   \
  } while (0)

#define DY_CUDA(call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return ::dy::fail(DY_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__),  \
                        __FILE__, __LINE__);                                                   \
  } while (0)

inline int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(DY_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return DY_OK;
}

int num_sms();   // SM count of the current device (cached)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---- device helpers -------------------------------------------------------------------------
__device__ __forceinline__ float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);   // a -> low half, b -> high half
  return *reinterpret_cast<uint32_t*>(&t);
}
// Identity the optimiser cannot see through: keeps a loop invariant in a register.
__device__ __forceinline__ uint32_t opaque(uint32_t x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ int opaque(int x) { asm volatile("" : "+r"(x)); return x; }
__device__ __forceinline__ float tanh_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}
// SiLU with one MUFU: x*sigmoid(x) = h + h*tanh(h), h = x/2
__device__ __forceinline__ float silu_fast(float x) {
  float h = 0.5f * x, t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

}  // namespace dy
