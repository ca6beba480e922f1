// Development microbenchmark: MUFU ex2 throughput per SM for f32 / f16x2 / bf16x2 operands as a function of resident warps.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu ; ./mufu_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* clk, int iters) {
  float a[16];
  uint32_t h[16];
  for (int i = 0; i < 16; ++i) {
    a[i] = -0.001f * (threadIdx.x + i);
    h[i] = 0xb800b900u + i;   // two small negative halves / bf16s
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 16; ++i) s += a[i] + (float)h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

int main() {
  float* out;
  long long* clk;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  const int iters = 2000;
  const char* names[3] = {"ex2.f32", "ex2.f16x2", "ex2.bf16x2"};
  for (int mode = 0; mode < 3; ++mode)
    for (int warps = 4; warps <= 32; warps *= 2) {
      if (mode == 0) k<0><<<148, warps * 32>>>(out, clk, iters);
      if (mode == 1) k<1><<<148, warps * 32>>>(out, clk, iters);
      if (mode == 2) k<2><<<148, warps * 32>>>(out, clk, iters);
      cudaDeviceSynchronize();
      long long c;
      cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
      const double instr = (double)iters * 16 * warps * 32;          // lane-instructions per SM
      const double per = (mode == 0 ? 1.0 : 2.0);
      printf("%-11s warps/SM=%2d  %8lld clk  %6.2f lane-instr/clk/SM  %6.2f exps/clk/SM  (%s)\n", names[mode], warps, c, instr / c,
             instr * per / c, cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
