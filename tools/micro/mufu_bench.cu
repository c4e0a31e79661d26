// MUFU throughput probe: lanes/clk/SM of the transcendental ops a SiLU epilogue could use.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/micro/mufu_bench.cu -o gpurun_out/mufu_bench && gpurun_out/mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 6) asm volatile("{.reg .f32 t; fma.rn.f32 t, %1, %1, %1; mov.b32 %0, t;}" : "=r"(y) : "r"(x));
  return y;
}

template <int OP>
__global__ void k(uint32_t* out, int iters, long long* cycles) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0x3c003c00u + threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = op<OP>(a[i]);
  }
  long long t1 = clock64();
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int lanes_per_op) {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  for (int threads : {128, 256, 512, 1024}) {
    k<OP><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_clk = double(threads) * iters * 8 * lanes_per_op / double(c);
    printf("%-22s threads/SM %4d : %6.2f results/clk/SM\n", name, threads, per_clk);
  }
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("tanh.approx.f32", 1);
  run<1>("ex2.approx.ftz.f32", 1);
  run<2>("rcp.approx.ftz.f32", 1);
  run<3>("tanh.approx.f16x2", 2);
  run<4>("tanh.approx.bf16x2", 2);
  run<5>("ex2.approx.f16x2", 2);
  run<6>("fma.rn.f32", 1);
  return 0;
}
