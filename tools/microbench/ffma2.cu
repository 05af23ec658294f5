// Diagnostic microbenchmark (not part of the product): issue throughput of FFMA vs FFMA2 (fma.rn.f32x2) on sm_100a.
#include <cuda_runtime.h>
#include <stdio.h>
__device__ __forceinline__ void ffma2(float& dx, float& dy, float ax, float ay, float bx, float by) {
  unsigned long long ra, rb, rc, rd;
  asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(ax), "f"(ay));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(bx), "f"(by));
  asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(dx), "f"(dy));
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(dx), "=f"(dy) : "l"(rd));
}
template <int MODE> __global__ void __launch_bounds__(256) k(float* out, float a, float b, int iters) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 1e-3f + i;
  float x = a + threadIdx.x * 1e-6f, y = b;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], x, y);
      } else {
#pragma unroll
        for (int i = 0; i < 16; i += 2) ffma2(acc[i], acc[i + 1], acc[i], acc[i + 1], x, x + 1.0f);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 4096;
  for (int mode = 0; mode < 2; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<148 * 8, 256>>>(d, 0.999f, 0.001f, iters); else k<1><<<148 * 8, 256>>>(d, 0.999f, 0.001f, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double fma = (double)148 * 8 * 256 * iters * 8 * 16;  // scalar FMAs
      printf("mode %s rep %d: %.3f ms  %.2f TFMA/s (=%.1f TFLOP/s)  warp-instr/clk/SM at 1.965GHz: %.2f\n", mode ? "FFMA2" : "FFMA ", rep, ms,
             fma / ms / 1e9, 2 * fma / ms / 1e9, (fma / 32 / (mode ? 2 : 1)) / (ms * 1e-3) / 1.965e9 / 148);
    }
  }
  return 0;
}
