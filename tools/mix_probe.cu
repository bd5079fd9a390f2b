// Do FP64-chain warps (deep rounds) and integer-heavy warps (sampler batches) overlap on one SM
// sub-partition?  8 warps per sub-partition; even warps run `rounds` deep rounds, odd warps run
// `batches` sampler batches.  Times: each group alone (the other group exits at once), then both.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o mix_probe mix_probe.cu
#include <cstdio>
#include "../cudabrot_b200/csrc/buddha_kernels.cuh"
using namespace buddha;

__global__ void __launch_bounds__(256) k(RenderParams p, double *out, unsigned *outi, int rounds, int batches) {
  __shared__ double2 sm[8][64];
  const int warp = threadIdx.x >> 5;
  double acc = 0; unsigned acci = 0;
  if ((warp >> 2) & 1) {   // warps 4..7 of the CTA: one per sub-partition (warp % 4), FP64 chains
    double cx = -0.2 + threadIdx.x * 1e-4, cy = 0.3, x = cx, y = cy;
#pragma unroll 1
    for (int i = 0; i < rounds; i++) {
#pragma unroll
      for (int u = 0; u < 24; u++) zstep<false>(x, y, cx, cy);
      if (!(norm4(x, y) <= 16.0)) { x = cx; y = cy; }
    }
    acc = x + y;
  } else {                 // warps 0..3: sampler batches
    unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned t1n = 0;
#pragma unroll 1
    for (int i = 0; i < batches; i++) {
      uint4 r = philox4x32_10(s, p);
      double gx = coord2_from_words(r.x, r.y), gy = coord2_from_words(r.z, r.w);
      const bool rej = rejected2(gx, gy);
      double a = gx, b = gy;
      zstep<false>(a, b, gx, gy);
      const bool in1 = !rej && !(norm4(a, b) > 16.0);
      zstep<false>(a, b, gx, gy);
      const bool in2 = in1 && !(norm4(a, b) > 16.0);
      acci += rej ? 1u : 0u; acci += in1 ? 1u : 0u;
      s += (unsigned long long)gridDim.x * blockDim.x;
      unsigned m = __ballot_sync(kFull, in2);
      if (in2) sm[warp][(t1n + __popc(m & lanemask_lt())) & 63] = make_double2(gx, gy);
      t1n += __popc(m);
    }
    acci += t1n;
    __syncwarp();
    acc = sm[warp][threadIdx.x & 31].x;
  }
  out[(blockIdx.x * blockDim.x + threadIdx.x) & 8191] = acc;
  outi[(blockIdx.x * blockDim.x + threadIdx.x) & 8191] = acci;
}

int main() {
  RenderParams p = {};
  for (int i = 0; i < 10; i++) { p.key0[i] = 1337u + 0x9E3779B9u * i; p.key1[i] = 0xBB67AE85u * i; }
  double *o; unsigned *oi; cudaMalloc(&o, 8 * 8192); cudaMalloc(&oi, 4 * 8192);
  cudaMemset(o, 0, 8 * 8192);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  auto run = [&](int ctas_per_sm, int rounds, int batches) {
    float ms = 0;
    for (int rep = 0; rep < 2; rep++) {
      cudaEventRecord(a); k<<<148 * ctas_per_sm, 256>>>(p, o, oi, rounds, batches);
      cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
    }
    return ms;
  };
  for (int cps : {1, 2, 3, 4}) {   // 8, 16, 24, 32 warps per SM: half FP64 warps, half sampler warps
    const int R = 4000;
    float tr = run(cps, R, 0);
    // choose batches so that the sampler group alone takes about as long
    float tb1 = run(cps, 0, R);
    int B = (int)(R * tr / tb1);
    float tb = run(cps, 0, B), both = run(cps, R, B);
    printf("%d warps/SM (%d FP64-chain + %d sampler per sub-partition): rounds alone %.2f ms, batches alone %.2f ms, together %.2f ms -> %.0f %% of the sum\n",
           cps * 8, cps, cps, tr, tb, both, 100.0 * both / (tr + tb));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
