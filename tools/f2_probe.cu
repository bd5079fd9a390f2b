// Does the packed FP32 instruction of sm_100 (FFMA2: fma.rn.f32x2) cost one issue slot or two?
// Cycles per warp and iteration for 32 FFMA, 32 FFMA2, 32 LOP3 and the pairwise mixes (8 independent
// chains, 8 warps per sub-partition): "mix ~ sum" = the two share the issue port cycle for cycle.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o f2_probe f2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int F1, int F2, int L3>   // instructions of each kind per unrolled slot
__global__ void __launch_bounds__(256) k(float *out, int n, float a, float b, unsigned m) {
  float v[8]; unsigned long long p[8]; unsigned w[8];
  for (int i = 0; i < 8; i++) { v[i] = threadIdx.x + i; p[i] = (unsigned long long)(threadIdx.x + i) * 0x3f8000013f800001ull; w[i] = threadIdx.x * 7 + i + 1; }
  unsigned long long ab, bb;
  asm("mov.b64 %0, {%1, %1};" : "=l"(ab) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
#pragma unroll 1
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int u = 0; u < 32; u++) {
      if (F1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[u & 7]) : "f"(a), "f"(b));
      if (F2) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[u & 7]) : "l"(ab), "l"(bb));
      if (L3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(w[u & 7]) : "r"(w[(u + 3) & 7]), "r"(m));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += v[i] + (float)(p[i] & 0xffff) + (float)w[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int F1, int F2, int L3> float run(float *o, int n) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a); k<F1, F2, L3><<<148 * 4, 256>>>(o, n, 1.0000001f, 1e-9f, 0x9E3779B9u);
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
  }
  return ms * 1e-3f * 1.965e9f / n / 8.0f;   // cycles per iteration and warp (8 warps per sub-partition)
}
int main() {
  float *o; cudaMalloc(&o, 4 * 148 * 4 * 256);
  const int n = 4000;
  printf("32 FFMA          : %.1f cycles per iteration and warp\n", run<1, 0, 0>(o, n));
  printf("32 FFMA2         : %.1f\n", run<0, 1, 0>(o, n));
  printf("32 LOP3          : %.1f\n", run<0, 0, 1>(o, n));
  printf("32 FFMA + 32 LOP3 : %.1f\n", run<1, 0, 1>(o, n));
  printf("32 FFMA2 + 32 LOP3: %.1f\n", run<0, 1, 1>(o, n));
  printf("32 FFMA + 32 FFMA2: %.1f\n", run<1, 1, 0>(o, n));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
