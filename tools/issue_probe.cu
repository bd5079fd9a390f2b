// Can the SM sub-partition issue integer work in the shadow of FP64 instructions?  Each loop
// iteration runs ND independent DFMAs (8 chains) and NI independent integer ops (8 chains), 32
// warps per SM.  Prints cycles per iteration per sub-partition warp: 2*ND + NI means an FP64
// instruction holds the issue port for both of its pipe cycles, max(2*ND, ND + NI) means it does not.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o issue_probe issue_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NI, int KIND>
__global__ void __launch_bounds__(256) k(double *out, unsigned *outi, int n, double a, double b, unsigned m) {
  double v[8]; unsigned w[8];
  for (int i = 0; i < 8; i++) { v[i] = threadIdx.x + i; w[i] = threadIdx.x * 7 + i; }
#pragma unroll 1
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int u = 0; u < (ND > NI ? ND : NI); u++) {
      if (u < ND) v[u & 7] = __fma_rn(v[u & 7], a, b);
      if (u < NI) {
        if (KIND == 0) w[u & 7] = (w[u & 7] ^ m) + (w[(u + 1) & 7] & m);      // LOP3 / IADD3 (alu)
        if (KIND == 1) w[u & 7] = w[u & 7] * m + w[(u + 1) & 7];               // IMAD (fma pipe)
        if (KIND == 2) { unsigned long long p = (unsigned long long)w[u & 7] * 0xD2511F53u; w[u & 7] = (unsigned)p ^ (unsigned)(p >> 32); }  // IMAD.WIDE + LOP
      }
    }
  }
  double s = 0; unsigned t = 0;
  for (int i = 0; i < 8; i++) { s += v[i]; t += w[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int ND, int NI, int KIND> float run(double *o, unsigned *oi, int n) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a); k<ND, NI, KIND><<<148 * 4, 256>>>(o, oi, n, 1.0000001, 1e-9, 0x9E3779B9u);
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
  }
  return ms * 1e-3f * 1.965e9f / n / 8.0f;  // cycles per iteration per warp of a sub-partition (8 warps each)
}
int main() {
  double *o; unsigned *oi; cudaMalloc(&o, 8 * 148 * 4 * 256); cudaMalloc(&oi, 4 * 148 * 4 * 256);
  const int n = 4000;
  const char *kn[3] = {"LOP3+IADD3", "IMAD", "IMAD.WIDE+LOP"};
  printf("ND DFMA + NI int per iteration -> cycles per iteration per warp (8 warps per sub-partition)\n");
#define ROW(K) printf("%-14s ND=32: NI=0 %.1f  NI=16 %.1f  NI=32 %.1f  NI=64 %.1f | ND=0: NI=32 %.1f NI=64 %.1f | ND=16 NI=64 %.1f\n", kn[K], \
    run<32, 0, K>(o, oi, n), run<32, 16, K>(o, oi, n), run<32, 32, K>(o, oi, n), run<32, 64, K>(o, oi, n), run<0, 32, K>(o, oi, n), run<0, 64, K>(o, oi, n), run<16, 64, K>(o, oi, n));
  ROW(0) ROW(1) ROW(2)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
