// Latency probes on one warp: dependent DFMA chain, one z-step, clock64-timed.  nvcc -arch sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double *out, long long *t, int n, double a, double b) {
  double v = threadIdx.x * 1e-3 + 0.5, x = 0.1 + threadIdx.x * 1e-4, y = 0.2, cx = -0.2, cy = 0.3;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) v = __fma_rn(v, a, b);
  }
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) {
      double a4 = __dmul_rn(y, y), b4 = __fma_rn(x, x, -a4), yn = __fma_rn(x, y, cy);
      x = __fma_rn(b4, 0.5, cx); y = yn;
    }
  }
  long long t2 = clock64();
  float f = threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) f = __fmaf_rn(f, 1.0001f, 0.5f);
  }
  long long t3 = clock64();
  out[threadIdx.x] = v + x + y + f;
  if (threadIdx.x == 0) { t[0] = t1 - t0; t[1] = t2 - t1; t[2] = t3 - t2; }
}
int main() {
  double *o; long long *t, h[3];
  cudaMalloc(&o, 8 * 1024); cudaMalloc(&t, 24);
  for (int warps = 1; warps <= 8; warps *= 2) {
    k<<<1, 32 * warps>>>(o, t, 4096, 1.0000001, 1e-9);
    cudaMemcpy(h, t, 24, cudaMemcpyDeviceToHost);
    printf("warps/SM %d: dfma chain %.1f clk/op, zstep %.1f clk/step, ffma chain %.1f clk/op\n", warps,
           h[0] / 65536.0, h[1] / 65536.0, h[2] / 65536.0);
  }
  return 0;
}
