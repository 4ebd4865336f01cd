// microbenchmark: FP32 FMA throughput, scalar FFMA vs packed fma.rn.f32x2, on sm_100a.  nvcc -arch=sm_100a -O3 ffma2.cu -o ffma2
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int MODE> __global__ void k(float *out, int iters, float s) {
  float a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  if (MODE == 0) {
    for (int i = 0; i < iters; i++) {
      a0 = fmaf(a0, s, 1.0f); a1 = fmaf(a1, s, 1.0f); a2 = fmaf(a2, s, 1.0f); a3 = fmaf(a3, s, 1.0f);
      a4 = fmaf(a4, s, 1.0f); a5 = fmaf(a5, s, 1.0f); a6 = fmaf(a6, s, 1.0f); a7 = fmaf(a7, s, 1.0f);
    }
  } else {
    unsigned long long p0, p1, p2, p3, ss, one;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p0) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p1) : "f"(a2), "f"(a3));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p2) : "f"(a4), "f"(a5));
    asm("mov.b64 %0, {%1, %2};" : "=l"(p3) : "f"(a6), "f"(a7));
    asm("mov.b64 %0, {%1, %2};" : "=l"(ss) : "f"(s), "f"(s));
    asm("mov.b64 %0, {%1, %2};" : "=l"(one) : "f"(1.0f), "f"(1.0f));
    for (int i = 0; i < iters; i++) { p0 = fma2(p0, ss, one); p1 = fma2(p1, ss, one); p2 = fma2(p2, ss, one); p3 = fma2(p3, ss, one); }
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(p0));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a2), "=f"(a3) : "l"(p1));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a4), "=f"(a5) : "l"(p2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a6), "=f"(a7) : "l"(p3));
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
  float *out; cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, grid = 148 * 8, block = 1024;
  for (int mode = 0; mode < 2; mode++) {
    for (int rep = 0; rep < 3; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<grid, block>>>(out, iters, 0.999f); else k<1><<<grid, block>>>(out, iters, 0.999f);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double fmas = double(grid) * block * iters * 8;
      printf("%s: %.3f ms, %.1f TFMA/s (%.1f TFLOP/s)\n", mode ? "fma.rn.f32x2" : "fma.rn.f32  ", ms, fmas / ms / 1e9, 2 * fmas / ms / 1e9);
    }
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
