// Does ptxas interleave an independent integer-heavy batch (Philox + coordinates + cardioid test +
// two tested steps = the sampler batch) into the latency shadow of a 24-step FP64 z-chain (a deep
// round) when both sit in one basic block?  Three kernels over the same grid: round only, batch
// only, both fused; time per loop iteration at several occupancies.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o fuse_probe fuse_probe.cu
#include <cstdio>
#include "../cudabrot_b200/csrc/buddha_kernels.cuh"
using namespace buddha;

__device__ __forceinline__ void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                             const RenderParams &p, int r) {
  uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
  uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
  uint32_t n0 = hi1 ^ c1 ^ p.key0[r];
  uint32_t n2 = hi0 ^ c3 ^ p.key1[r];
  c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

// volatile inline PTX: NVVM keeps volatile asm statements in source order
__device__ __forceinline__ void zstep_v(double &x, double &y, double cx, double cy) {
  double a4, b4, yn;
  asm volatile("mul.rn.f64 %0, %1, %1;" : "=d"(a4) : "d"(y));
  asm volatile("neg.f64 %0, %0;" : "+d"(a4));
  asm volatile("fma.rn.f64 %0, %1, %1, %2;" : "=d"(b4) : "d"(x), "d"(a4));
  asm volatile("fma.rn.f64 %0, %1, %2, %3;" : "=d"(yn) : "d"(x), "d"(y), "d"(cy));
  asm volatile("fma.rn.f64 %0, %1, 0d3FE0000000000000, %2;" : "=d"(x) : "d"(b4), "d"(cx));
  y = yn;
}
__device__ __forceinline__ void philox_round_v(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                               uint32_t k0, uint32_t k1) {
  uint32_t hi0, lo0, hi1, lo1, n0, n2;
  asm volatile("mul.hi.u32 %0, %2, 0xD2511F53; mul.lo.u32 %1, %2, 0xD2511F53;" : "=r"(hi0), "=r"(lo0) : "r"(c0));
  asm volatile("mul.hi.u32 %0, %2, 0xCD9E8D57; mul.lo.u32 %1, %2, 0xCD9E8D57;" : "=r"(hi1), "=r"(lo1) : "r"(c2));
  asm volatile("xor.b32 %0, %1, %2; xor.b32 %0, %0, %3;" : "=r"(n0) : "r"(hi1), "r"(c1), "r"(k0));
  asm volatile("xor.b32 %0, %1, %2; xor.b32 %0, %0, %3;" : "=r"(n2) : "r"(hi0), "r"(c3), "r"(k1));
  c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
}

__global__ void __launch_bounds__(256) k5(RenderParams p, double *out, unsigned *outi, int n) {
  __shared__ double2 sm[8][64];
  double cx = -0.2 + threadIdx.x * 1e-4, cy = 0.3, x = cx, y = cy;
  unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned acc = 0, t1n = 0;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    uint32_t c0 = (uint32_t)s, c1 = (uint32_t)(s >> 32), c2 = 0u, c3 = 0u;
    double gx = 0, gy = 0, a = 0, b = 0;
    double i2 = 0, q0 = 0, q = 0, sx = 0, lhs = 0, t = 0, bb = 0;
    bool rej = false, in1 = false, in2 = false;
#pragma unroll
    for (int u = 0; u < 24; u++) {
      zstep_v(x, y, cx, cy);
      if (u < 10) philox_round_v(c0, c1, c2, c3, p.key0[u], p.key1[u]);
      if (u == 10) gx = coord2_from_words(c0, c1);
      if (u == 11) gy = coord2_from_words(c2, c3);
      if (u == 12) { i2 = __dmul_rn(gy, gy); q0 = __dadd_rn(gx, -0.5); t = __dadd_rn(gx, 2.0); }
      if (u == 13) { q = __fma_rn(q0, q0, i2); bb = __fma_rn(t, t, i2); }
      if (u == 14) { sx = __fma_rn(q0, 2.0, q); }
      if (u == 15) { lhs = __dmul_rn(q, sx); rej = (lhs < i2) || (bb < 0.25); a = gx; b = gy; }
      if (u == 16) zstep_v(a, b, gx, gy);
      if (u == 18) in1 = !rej && !(norm4(a, b) > 16.0);
      if (u == 19) zstep_v(a, b, gx, gy);
      if (u == 21) in2 = in1 && !(norm4(a, b) > 16.0);
    }
    acc += rej ? 1u : 0u;
    acc += in1 ? 1u : 0u;
    s += (unsigned long long)gridDim.x * blockDim.x;
    if (!(norm4(x, y) <= 16.0)) { x = cx; y = cy; }
    unsigned m = __ballot_sync(kFull, in2);
    if (in2) { sm[threadIdx.x >> 5][(t1n + __popc(m & lanemask_lt())) & 63] = make_double2(gx, gy); }
    t1n += __popc(m);
  }
  out[s & 1023] = x + y + sm[threadIdx.x >> 5][threadIdx.x & 31].x;
  outi[s & 1023] = acc + t1n;
}

// mode 4: the batch sliced by hand between the z-steps of the round
__global__ void __launch_bounds__(256) k4(RenderParams p, double *out, unsigned *outi, int n) {
  __shared__ double2 sm[8][64];
  double cx = -0.2 + threadIdx.x * 1e-4, cy = 0.3, x = cx, y = cy;
  unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned acc = 0, t1n = 0;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    uint32_t c0 = (uint32_t)s, c1 = (uint32_t)(s >> 32), c2 = 0u, c3 = 0u;
    double gx = 0, gy = 0, a = 0, b = 0;
    double i2 = 0, q0 = 0, q = 0, sx = 0, lhs = 0, t = 0, bb = 0;
    bool rej = false, in1 = false, in2 = false;
#pragma unroll
    for (int u = 0; u < 24; u++) {
      zstep<false>(x, y, cx, cy);
      if (u < 10) philox_round(c0, c1, c2, c3, p, u);
      if (u == 10) gx = coord2_from_words(c0, c1);
      if (u == 11) gy = coord2_from_words(c2, c3);
      if (u == 12) { i2 = __dmul_rn(gy, gy); q0 = __dadd_rn(gx, -0.5); t = __dadd_rn(gx, 2.0); }
      if (u == 13) { q = __fma_rn(q0, q0, i2); bb = __fma_rn(t, t, i2); }
      if (u == 14) { sx = __fma_rn(q0, 2.0, q); }
      if (u == 15) { lhs = __dmul_rn(q, sx); rej = (lhs < i2) || (bb < 0.25); a = gx; b = gy; }
      if (u == 16) zstep<false>(a, b, gx, gy);
      if (u == 18) in1 = !rej && !(norm4(a, b) > 16.0);
      if (u == 19) zstep<false>(a, b, gx, gy);
      if (u == 21) in2 = in1 && !(norm4(a, b) > 16.0);
    }
    acc += rej ? 1u : 0u;
    acc += in1 ? 1u : 0u;
    s += (unsigned long long)gridDim.x * blockDim.x;
    if (!(norm4(x, y) <= 16.0)) { x = cx; y = cy; }
    unsigned m = __ballot_sync(kFull, in2);
    if (in2) { sm[threadIdx.x >> 5][(t1n + __popc(m & lanemask_lt())) & 63] = make_double2(gx, gy); }
    t1n += __popc(m);
  }
  out[s & 1023] = x + y + sm[threadIdx.x >> 5][threadIdx.x & 31].x;
  outi[s & 1023] = acc + t1n;
}

template <int kMode>  // 1 = round, 2 = batch, 3 = fused
__global__ void __launch_bounds__(256) k(RenderParams p, double *out, unsigned *outi, int n) {
  __shared__ double2 sm[8][64];
  double cx = -0.2 + threadIdx.x * 1e-4, cy = 0.3, x = cx, y = cy;
  unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned acc = 0, t1n = 0;
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    bool in2 = false;
    double gx = 0, gy = 0;
    if (kMode & 2) {
      uint4 r = philox4x32_10(s, p);
      gx = coord2_from_words(r.x, r.y);
      gy = coord2_from_words(r.z, r.w);
      const bool rej = rejected2(gx, gy);
      double a = gx, b = gy;
      zstep<false>(a, b, gx, gy);
      const bool in1 = !rej && !(norm4(a, b) > 16.0);
      zstep<false>(a, b, gx, gy);
      in2 = in1 && !(norm4(a, b) > 16.0);
      acc += rej ? 1u : 0u;
      acc += in1 ? 1u : 0u;
      s += (unsigned long long)gridDim.x * blockDim.x;
    }
    if (kMode & 1) {
#pragma unroll
      for (int u = 0; u < 24; u++) zstep<false>(x, y, cx, cy);
      if (!(norm4(x, y) <= 16.0)) { x = cx; y = cy; }
    }
    if (kMode & 2) {
      unsigned m = __ballot_sync(kFull, in2);
      if (in2) { sm[threadIdx.x >> 5][(t1n + __popc(m & lanemask_lt())) & 63] = make_double2(gx, gy); }
      t1n += __popc(m);
    }
  }
  out[s & 1023] = x + y + sm[threadIdx.x >> 5][threadIdx.x & 31].x;
  outi[s & 1023] = acc + t1n;
}

int main() {
  RenderParams p = {};
  for (int i = 0; i < 10; i++) { p.key0[i] = 1337u + 0x9E3779B9u * i; p.key1[i] = 0xBB67AE85u * i; }
  double *o; unsigned *oi;
  cudaMalloc(&o, 8 * 8192); cudaMalloc(&oi, 4 * 1024);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int n = 20000;
  for (int warps_per_sm : {8, 16, 24, 32, 40}) {
    int ctas = 148 * warps_per_sm / 8;  // 8-warp CTAs
    float ms[6] = {0, 0, 0, 0, 0, 0};
    for (int mode = 1; mode <= 5; mode++) {
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(a);
        if (mode == 1) k<1><<<ctas, 256>>>(p, o, oi, n);
        if (mode == 2) k<2><<<ctas, 256>>>(p, o, oi, n);
        if (mode == 3) k<3><<<ctas, 256>>>(p, o, oi, n);
        if (mode == 4) k4<<<ctas, 256>>>(p, o, oi, n);
        if (mode == 5) k5<<<ctas, 256>>>(p, o, oi, n);
        cudaEventRecord(b); cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms[mode], a, b);
      }
    }
    double clk = 1.965e9;
    printf("warps/SM %2d: cycles per iteration per SMSP-warp-slot: round %.0f  batch %.0f  fused %.0f  (sum %.0f)  hand-sliced %.0f  volatile-asm-sliced %.0f | per SM sub-partition issue-cycles per iteration*warps: fused %.1f\n",
           warps_per_sm, ms[1] * 1e-3 * clk / n, ms[2] * 1e-3 * clk / n, ms[3] * 1e-3 * clk / n,
           (ms[1] + ms[2]) * 1e-3 * clk / n, ms[4] * 1e-3 * clk / n, ms[5] * 1e-3 * clk / n, ms[5] * 1e-3 * clk / n / (warps_per_sm / 4.0));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
