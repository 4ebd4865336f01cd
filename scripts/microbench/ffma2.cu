// microbenchmark: FP32 pipe rates on sm_100a — scalar 3-register FFMA / FADD / FMNMX against the packed f32x2 forms, alone and
// interleaved.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 ffma2.cu -o ffma2
// Prints warp-instructions per clock per SM (4 SMSPs; 1 inst/clk/SMSP issue limit => 4.0 is the ceiling) from %clock64.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ float fmas(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }
__device__ __forceinline__ float adds(float a, float b) { float d; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float mins(float a, float b) { float d; asm volatile("min.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ u64 pack(float a, float b) { u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(a), "f"(b)); return d; }
__device__ __forceinline__ float lo(u64 p) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p)); return a + b; }

// MODE: 0 FFMA  1 FFMA2  2 FADD  3 FADD2  4 FMNMX  5 FFMA+FMNMX (1:1)  6 FFMA2+FMNMX (1:1)  7 FMUL2  8 FFMA+FADD (1:1)
template <int MODE> __global__ void k(float *out, const float *in, int iters, long long *cycles) {
  const float s = in[0], c = in[1]; // register operands (not immediates / constant bank)
  float a[8];
  u64 p[8];
#pragma unroll
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; p[i] = pack(a[i], a[i] + 0.5f); }
  const u64 ss = pack(s, s), cc = pack(c, c);
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = fmas(a[i], s, c);
      if (MODE == 1) p[i] = fma2(p[i], ss, cc);
      if (MODE == 2) a[i] = adds(a[i], c);
      if (MODE == 3) p[i] = add2(p[i], cc);
      if (MODE == 4) a[i] = mins(a[i], c);
      if (MODE == 5) { if (i & 1) a[i] = mins(a[i], c); else a[i] = fmas(a[i], s, c); }
      if (MODE == 6) { if (i & 1) a[i] = mins(a[i], c); else p[i] = fma2(p[i], ss, cc); }
      if (MODE == 7) p[i] = mul2(p[i], ss);
      if (MODE == 8) { if (i & 1) a[i] = adds(a[i], c); else a[i] = fmas(a[i], s, c); }
    }
  }
  const long long t1 = clock64();
  float r = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) r += a[i] + lo(p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int MODE> void run(const char *name, float *out, float *in, long long *cyc) {
  const int iters = 4096, grid = 148, block = 1024; // 32 warps / SM = 8 per SMSP
  k<MODE><<<grid, block>>>(out, in, iters, cyc);
  k<MODE><<<grid, block>>>(out, in, iters, cyc);
  cudaDeviceSynchronize();
  long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double warpInst = 32.0 * iters * 8; // per SM
  printf("%-22s %8lld cycles  %.3f warp-inst/clk/SM\n", name, c, warpInst / double(c));
}
int main() {
  float *out, *in; long long *cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&in, 8); cudaMalloc(&cyc, 8);
  const float h[2] = {0.999f, 1.0f}; cudaMemcpy(in, h, 8, cudaMemcpyHostToDevice);
  run<0>("FFMA (3 reg)", out, in, cyc);
  run<1>("FFMA2 (fma.f32x2)", out, in, cyc);
  run<2>("FADD", out, in, cyc);
  run<3>("FADD2 (add.f32x2)", out, in, cyc);
  run<7>("FMUL2 (mul.f32x2)", out, in, cyc);
  run<4>("FMNMX", out, in, cyc);
  run<5>("FFMA + FMNMX 1:1", out, in, cyc);
  run<6>("FFMA2 + FMNMX 1:1", out, in, cyc);
  run<8>("FFMA + FADD 1:1", out, in, cyc);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
