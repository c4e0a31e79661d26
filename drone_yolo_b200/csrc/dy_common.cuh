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

int num_sms();   // SM count of the current device (cached per device)

// The dynamic shared-memory opt-in (cudaFuncAttributeMaxDynamicSharedMemorySize) is a per-DEVICE attribute of a kernel:
// `seen` holds one bit per device ordinal; returns true the first time the calling thread's current device asks.
inline bool first_use_on_device(unsigned long long* seen) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
  const unsigned long long bit = 1ull << dev;
  if (*seen & bit) return false;
  *seen |= bit;
  return true;
}

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
// ---- Detect decode math, shared by decode.cu and the fused Detect tail of conv_igemm.cu (same source -> same results) ----
static constexpr int kRegMax = 16;
// Expectation of a 16-bin distribution given its logits (DFL.forward, nn/modules/block.py:73-76: softmax over the bins,
// weights 0..15).
__device__ __forceinline__ float dfl_expect(const float (&x)[kRegMax]) {
  float m = x[0];
#pragma unroll
  for (int i = 1; i < kRegMax; ++i) m = fmaxf(m, x[i]);
  float s = 0.f, t = 0.f;
#pragma unroll
  for (int i = 0; i < kRegMax; ++i) {
    const float e = __expf(__fsub_rn(x[i], m));
    s = __fadd_rn(s, e);
    t = fmaf(static_cast<float>(i), e, t);
  }
  return __fdividef(t, s);
}
// dist2bbox(xywh=True) * stride (utils/tal.py:348-357, head.py:129): lt, rb = chunk; x1y1 = a - lt; x2y2 = a + rb;
// c = (x1y1 + x2y2) / 2; wh = x2y2 - x1y1.  Explicit roundings: no FMA contraction, identical in every caller.
__device__ __forceinline__ void dist2bbox_xywh(float ax, float ay, const float (&d)[4], float stride, float (&b)[4]) {
  const float x1 = __fsub_rn(ax, d[0]), y1 = __fsub_rn(ay, d[1]), x2 = __fadd_rn(ax, d[2]), y2 = __fadd_rn(ay, d[3]);
  b[0] = __fmul_rn(__fmul_rn(__fadd_rn(x1, x2), 0.5f), stride);
  b[1] = __fmul_rn(__fmul_rn(__fadd_rn(y1, y2), 0.5f), stride);
  b[2] = __fmul_rn(__fsub_rn(x2, x1), stride);
  b[3] = __fmul_rn(__fsub_rn(y2, y1), stride);
}
__device__ __forceinline__ float sigmoid_fast(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

}  // namespace dy
