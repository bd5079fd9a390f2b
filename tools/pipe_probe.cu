// Which instructions share an issue/dispatch port with the FP64 pipe on B200?  For each integer /
// FP32 instruction X: cycles per instruction alone (8 independent chains, 8 warps per
// sub-partition), and the time of 32 X + 32 DFMA together.  "together ~ alone_X + 64" = X and DFMA
// are additive (same port); "together ~ max" = they overlap.
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define OPS(X) \
  X(LOP3,   "xor.b32 %0, %0, %1;",                      1) \
  X(IADD,   "add.u32 %0, %0, %1;",                      1) \
  X(IMAD,   "mad.lo.u32 %0, %0, %2, %1;",               1) \
  X(IMADHI, "mul.hi.u32 %0, %0, %2;",                   1) \
  X(LOP3ABC, "lop3.b32 %0, %0, %1, %2, 0x96;",           1) \
  X(FSEL, "{ .reg .pred p; .reg .f32 f, g; setp.lt.u32 p, %1, %2; mov.b32 f, %0; mov.b32 g, %1; selp.f32 f, f, g, p; mov.b32 %0, f; }", 2) \
  X(IMADMOV, "mad.lo.u32 %0, %1, 1, 0;", 1) \
    X(IMADWIDE2, "{ .reg .b64 t; .reg .b32 lo, hi; mul.wide.u32 t, %0, %2; mov.b64 {lo, hi}, t; xor.b32 %0, lo, hi; }", 2) \
  X(HI_LO, "{ .reg .b32 lo, hi; mul.hi.u32 hi, %0, %2; mul.lo.u32 lo, %0, %1; xor.b32 %0, lo, hi; }", 3) \
  X(SHF,    "shf.l.wrap.b32 %0, %0, %0, 7;",            1) \
  X(ISETP_SEL, "{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %1, %0, p; add.u32 %0, %0, 1; }", 3) \
  X(POPC,   "popc.b32 %0, %0;",                         1) \
  X(PRMT,   "prmt.b32 %0, %0, %1, 0x3201;",             1) \
  X(FFMA,   "{ .reg .f32 f; mov.b32 f, %0; fma.rn.f32 f, f, 0f3F800001, 0f3A83126F; mov.b32 %0, f; }", 1) \
  X(VOTE,   "{ .reg .pred p; setp.ne.u32 p, %0, 0; vote.sync.ballot.b32 %0, p, 0xffffffff; }", 2) \
  X(I2F64,  "{ .reg .f64 d; .reg .b64 u; cvt.u64.u32 u, %0; cvt.rn.f64.u64 d, u; mov.b64 {%0, _}, d; }", 1)

#define ENUM(NAME, STR, CNT) k##NAME,
enum { OPS(ENUM) kNumOps };
template <int OP, bool WITH_D>
__global__ void __launch_bounds__(256) k(double *out, unsigned *outi, int n, double a, double b, unsigned m) {
  double v[8]; unsigned w[8];
  for (int i = 0; i < 8; i++) { v[i] = threadIdx.x + i; w[i] = threadIdx.x * 7 + i + 1; }
#pragma unroll 1
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int u = 0; u < 32; u++) {
      if (WITH_D) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[u & 7]) : "d"(a), "d"(b));
#define CASE(NAME, STR, CNT) if (OP == k##NAME) asm volatile(STR : "+r"(w[u & 7]) : "r"(w[(u + 3) & 7]), "r"(m));
      OPS(CASE)
    }
  }
  double s = 0; unsigned t = 0;
  for (int i = 0; i < 8; i++) { s += v[i]; t += w[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s; outi[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int OP, bool WITH_D> float run(double *o, unsigned *oi, int n) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float ms = 0;
  for (int rep = 0; rep < 2; rep++) {
    cudaEventRecord(a); k<OP, WITH_D><<<148 * 4, 256>>>(o, oi, n, 1.0000001, 1e-9, 0x9E3779B9u);
    cudaEventRecord(b); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
  }
  return ms * 1e-3f * 1.965e9f / n / 8.0f;
}
int main() {
  double *o; unsigned *oi; cudaMalloc(&o, 8 * 148 * 4 * 256); cudaMalloc(&oi, 4 * 148 * 4 * 256);
  const int n = 4000;
  printf("32 DFMA alone: %.1f cycles per iteration per warp (8 warps per sub-partition)\n", run<-1, true>(o, oi, n));
  int op = 0;
#define ROW(NAME, STR, CNT) { float al = run<k##NAME, false>(o, oi, n); float wd = run<k##NAME, true>(o, oi, n); \
    printf("%-10s (%d SASS-ish instr): 32 alone %.1f (%.2f per instr)   with 32 DFMA %.1f   -> %s\n", #NAME, CNT, al, al / 32 / CNT, wd, \
           wd > al + 48 ? "ADDITIVE with FP64" : (wd < (al > 64 ? al : 64) + 16 ? "overlaps FP64" : "partial")); op++; }
  OPS(ROW)
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
