// buddha_kernels.cuh -- sm_100a device code of libbuddha: the Buddhabrot hot path.
//
// Replaces the reference's two kernels (InitializeRNG cudabrot.cu:146-149, DrawBuddhabrot
// :379-414 with :284-365) and moves its serial host tone-map (:416-468) to the GPU.
// Results are bit-identical to the reference's arithmetic: every FP64 operation
// is pinned with a round-to-nearest intrinsic so neither NVVM nor ptxas can re-associate or
// re-contract it.  DESIGN.md derives the identities used below.
//
// Orbit state is kept SCALED BY TWO: X = 2*re, Y = 2*im, CX = 2*c_re, CY = 2*c_im.  Scaling by a
// power of two is exact (no value on this path can overflow or go subnormal),
// so each rounding below is the reference's rounding of the same quantity times a power of two:
//     A4 = rn(Y*Y)            = 4*rn(im*im)                 DMUL
//     B4 = fma(X, X, -A4)     = 4*fma(re, re, -rn(im*im))   DFMA
//     X' = fma(B4, 0.5, CX)   = 2*rn(c_re + B)              DFMA  (replaces DADD re+re AND DADD c+t2)
//     Y' = fma(X, Y, CY)      = 2*fma(rn(re+re), im, c_im)  DFMA
//     S4 = fma(Y',Y', rn(X'*X')) = 4*fma(im',im', rn(re'*re'));  escaped <=> S4 > 16
// i.e. 4 FP64 instructions per step instead of the reference's 5, with identical bits.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace buddha {

// Occupancy: 3 CTAs x 8 warps per SM (24 warps, <= 80 registers, 8176 B of work stacks per warp).
// Measured (tools/gpu_ab.sh, profiles/r02_occupancy_ab.txt): 22, 24, 26 and 27 warps per SM are within
// 2 % of each other -- the kernel is bound by instruction issue, not by latency -- and 3 x 8 leaves
// the registers the co-resident apply / drain CTAs of the tiled pipeline need.
// BUDDHA_WARPS_PER_CTA / BUDDHA_CTAS_PER_SM: A/B builds.
#ifndef BUDDHA_WARPS_PER_CTA
#define BUDDHA_WARPS_PER_CTA 8
#endif
#ifndef BUDDHA_CTAS_PER_SM
#define BUDDHA_CTAS_PER_SM 3
#endif
constexpr int kWarpsPerCta = BUDDHA_WARPS_PER_CTA;
constexpr int kCtasPerSm = BUDDHA_CTAS_PER_SM;
constexpr int kThreadsPerCta = kWarpsPerCta * 32;
// Work stacks (entries).  A phase pops at most 32 entries and pushes at most 32 (the sampler: 64
// into t0), and runs only while its targets hold < 32, so no stack exceeds 63 (t0: 95).  `late` and
// `orb` share one array and grow towards each other, and so do `t0` and `t2`: t0 + t2 <= 95 + 31
// (the sampler runs only while t2 < 32, the first tier leaves t0 <= 63 when it fills t2 to <= 63),
// and late + orb <= kZJoint + 32
// because every phase that pushes to them starts only while late + orb <= kZJoint (else `late`
// runs first, with whatever it holds).
constexpr int kDeepCap = 63;
constexpr int kZCap = 87, kZJoint = kZCap - 32;
// The sampler draws two candidates per lane and loop trip: twice the instruction-level parallelism
// in the Philox chains (ten dependent multiply-xor rounds), one loop trip and one stack update for
// both.  Config 1 +3.6 %, config 2 +3.3 %, config 4 +2.2 % (profiles/r02_sampler_two_per_lane_ab.txt;
// BUDDHA_GEN2=0: one per lane, A/B builds).
#ifndef BUDDHA_GEN2
#define BUDDHA_GEN2 1
#endif
constexpr bool kGen2 = BUDDHA_GEN2 != 0;
constexpr int kCCap = kGen2 ? 126 : 94;        // (t0 then holds up to 31 + 64 while t2 holds <= 31)
constexpr int kSpillPerWarp = 32;                // a warp leaves the kernel with < 32 accepted samples
constexpr int kDrainLong = 1024;                 // leftovers this long get a warp each, if they are few
constexpr int kChunk = 4096;                   // granularity of launch sizes (host side)
constexpr int kMinChunk = 1024, kMaxChunk = 16384;  // sample indices a warp takes per cursor grab
#ifndef BUDDHA_T1_END
#define BUDDHA_T1_END 6
#endif
#ifndef BUDDHA_T2_END
#define BUDDHA_T2_END 22
#endif
constexpr int kT1End = BUDDHA_T1_END;          // the first exact tier covers steps 1 .. kT1End
constexpr int kT2End = BUDDHA_T2_END;          // tier 2 covers steps kT1End+1 .. kT2End
constexpr int kLateSteps = 24;                 // per-step-tested steps per `late` batch
constexpr int kBlock = 24;                     // unchecked steps per deep round (= kLateSteps)
constexpr int kDeepExit = 31;                  // leave a phase when fewer lanes than this are busy
constexpr int kOrbExit = 28;                   // (tier ends, kBlock and the two exits: measured sweep,
                                               //  profiles/r01_summary.md)
constexpr unsigned kFull = 0xffffffffu;
constexpr int kMaxChannels = 4;                // fused multi-channel render
constexpr int kMaxBands = 2 * kMaxChannels - 1;  // distinct channel sets along the escape-step axis
constexpr int kOrbStepBits = 28;               // fused: orbit entries carry (band << 28) | steps

// Fast binning: T = fma(X, inv_half, C0) lands in [1.5*2^40, 1.5*2^40 + 2^20) for in-range
// quotients, where the low mantissa word holds quotient * 2^12.
constexpr int kBinFracBits = 12;
constexpr uint32_t kBinClearMask = 0xffcu;     // fraction bits that must not all be zero (>= 2^-10)
constexpr uint32_t kBinHiWord = 0x42780000u;   // high word of 1.5 * 2^40

enum CounterSlot {
  kCntRejected = 0, kCntHitMax, kCntTooEarly, kCntAccepted, kCntEscapeIters, kCntOrbitPoints,
  kCntIncrements, kCntExecuted, kCntShortcut, kCntExactBins, kCntSlots
};
// fused render: accumulators that follow the kCntSlots common ones, kChSlots per index; kChHit and
// kChOver are indexed by CHANNEL, the other three by BAND (see RenderParams)
enum ChannelSlot { kChHit = 0, kChOver, kChAccepted, kChPoints, kChIncrements, kChSlots };
constexpr int kChCount = (kMaxBands * kChSlots + 1) & ~1;   // accumulators per warp (even: 16-byte size)

struct RenderParams {
  // canvas in the reference's form (exact binning path), cudabrot.cu:46-58
  int32_t w, h;
  double min_re, min_im, delta_re, delta_im;
  // fast binning constants (host-computed, see buddha_api.cu: make_fast_bin)
  double inv_half_re, inv_half_im;
  double c0_re, c0_im;          // (NaN when the canvas does not admit the fast path: nothing is "in range")
  int32_t fast_bin;
  int32_t max_it, min_it;
  int32_t shortcut;
  int32_t ship;                 // burning-ship variant (only the simple kernel reads this at run time)
  // fused multi-channel render: channel k accepts a sample that escapes at step `it` iff
  // ch_min[k] <= it - 1 < ch_max[k].  The channels' windows cut the step axis into segments with a
  // constant set of accepting channels; each distinct non-empty set is a BAND with its own
  // histogram (hist + band * band_stride), so every orbit point costs ONE increment however many
  // channels take it, and channel k = the sum of the bands whose set contains k (formed when the
  // channel is read; sums are mod 2^32 like the cells themselves).  max_it / min_it above are the
  // largest ch_max / the smallest ch_min.
  int32_t n_ch;
  int32_t ch_max[kMaxChannels];
  int32_t ch_low;               // smallest ch_max: below it no channel has hit its limit yet
  int32_t n_bands, n_seg;
  int32_t seg_start[2 * kMaxChannels + 1];  // segment s covers steps seg_start[s] .. seg_start[s+1]-1
  int32_t seg_band[2 * kMaxChannels];       // its band, or -1 if no channel accepts there
  uint32_t band_stride;         // cells per band (= w * h)
  // Privatised copies for tiny canvases (north_star (d)): when every orbit point lands in a few
  // thousand cells, reductions to the same address serialise in L2 (64x64: 1.9e10 samples/s
  // against 1.0e11 from 256x256 up, profiles/r02_dense_canvas.txt).  CTA b then adds into copy
  // b % n_copies of the histogram (copy k at hist + k * copy_stride, all of them L2-resident) and
  // fold_copies_kernel sums them into copy 0 after every render call.  1 = off.
  uint32_t n_copies, copy_stride;
  uint32_t key0[10], key1[10];  // Philox round keys: key + r * (W0, W1)
  unsigned long long end;       // one past the last sample index of this launch
  uint32_t chunk;               // sample indices a warp takes per cursor grab (multiple of 32):
                                // large for few atomics and counter flushes, small enough that
                                // every warp of the grid gets several chunks
  // tile-binned scatter for histograms much larger than L2 (0 = off, see scatter())
  int32_t tile_shift;           // log2(cells per tile)
  int32_t n_tiles;
  uint32_t n_warps;             // warps of the full persistent grid = lists per tile
  uint32_t *tcount;             // [tile * n_warps + warp] appends ATTEMPTED so far (may exceed cap)
  const uint32_t *tile_cap;     // [tile] capacity of each of the tile's lists
  const uint32_t *tile_base;    // [tile] first pool entry of the tile; list w starts at + w * cap
  uint32_t *pool;               // tile-local cell offsets
  // certificate queues (cert_phase): cert_cap entries per warp of the persistent grid, 0 = off.
  // Pending samples grow down from the top of a warp's block, failed ones (waiting to return to
  // `deep`) up from its bottom.  meta = (last, age, period / state, -).
  double2 *cert_c, *cert_z;
  uint4 *cert_m;
  uint32_t cert_cap;
};

// ---- small helpers --------------------------------------------------------------------------

__device__ __forceinline__ void red_add_u32(uint32_t *addr) {
  asm volatile("red.global.add.u32 [%0], 1;" ::"l"(addr) : "memory");
}

// The same, predicated inside the instruction stream: no branch around it.
__device__ __forceinline__ void red_add_u32_if(uint32_t *addr, bool pred) {
  asm volatile("{ .reg .pred p; setp.ne.u32 p, %1, 0; @p red.global.add.u32 [%0], 1; }"
               ::"l"(addr), "r"((unsigned)pred) : "memory");
}

__device__ __forceinline__ unsigned lane_id() {
  unsigned l;
  asm("mov.u32 %0, %%laneid;" : "=r"(l));
  return l;
}

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// One step of z <- z^2 + c on the scaled state (4 FP64 instructions).  Ship step: the burning-ship
// variant (RENDER_BURNING_SHIP, cudabrot.cu:15-17, :327-330, :353-356), real = |real|, imag =
// |imag| before the step -- only the cross term sees the difference, as operand modifiers
// (the reference's own SASS for that build: DADD |re|,|re|; DFMA r2,|im|,c_im).
template <bool kShipStep>
__device__ __forceinline__ void zstep(double &x, double &y, double cx, double cy) {
  const double a4 = __dmul_rn(y, y);
  const double b4 = __fma_rn(x, x, -a4);
  const double yn = kShipStep ? __fma_rn(fabs(x), fabs(y), cy) : __fma_rn(x, y, cy);
  x = __fma_rn(b4, 0.5, cx);
  y = yn;
}
// Kernel variants (template parameter kVar): bit 0 = burning ship, bit 1 = fused multi-channel.
constexpr int kVarShip = 1, kVarFused = 2;
#define BUDDHA_ZSTEP(x, y, cx, cy) zstep<(kVar & kVarShip) != 0>((x), (y), (cx), (cy))

// 4 * (re^2 + im^2) with the reference's rounding order (cudabrot.cu:336).
__device__ __forceinline__ double norm4(double x, double y) {
  return __fma_rn(y, y, __dmul_rn(x, x));
}

// Philox4x32-10 (curand_philox4x32_x.h:88-91,159-192), counter = (s_lo, s_hi, 0, 0).
__device__ __forceinline__ uint4 philox4x32_10(unsigned long long s, const RenderParams &p) {
  uint32_t c0 = (uint32_t)s, c1 = (uint32_t)(s >> 32), c2 = 0u, c3 = 0u;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    // (one IMAD.WIDE each; splitting them into IMAD.HI + IMAD was measured 4..6 % slower)
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    uint32_t n0 = hi1 ^ c1 ^ p.key0[r];
    uint32_t n2 = hi0 ^ c3 ^ p.key1[r];
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
  }
  return make_uint4(c0, c1, c2, c3);
}

// _curand_uniform_double_hq (curand_uniform.h:101-105) followed by cudabrot.cu:392-393, scaled by
// two: returns 2*(u*4-2) = fma(u, 8, -4).  The 53-bit integer converts exactly (one I2F.F64.U64 on
// the XU pipe, which this kernel does not use otherwise).
__device__ __forceinline__ double coord2_from_words(uint32_t x, uint32_t y) {
  const unsigned long long z = ((unsigned long long)(y >> 11) << 32) | (x ^ (y << 21));
  double u = __fma_rn(__ull2double_rn(z), 0x1p-53, 0x1p-54);
  return __fma_rn(u, 8.0, -4.0);
}

// InMainCardioid || InOrder2Bulb (cudabrot.cu:284-298) on the scaled candidate:
// 4*i2, 2*q0, 4*q, 16*lhs vs 16*rn(i2*0.25) = 4*i2, and 4*b vs 4/16.
__device__ __forceinline__ bool rejected2(double cx, double cy) {
  double i2 = __dmul_rn(cy, cy);            // 4 * rn(im*im)
  double q0 = __dadd_rn(cx, -0.5);          // 2 * (re - 0.25)
  double q = __fma_rn(q0, q0, i2);          // 4 * q
  double s = __fma_rn(q0, 2.0, q);          // 4 * rn(q0 + q)
  double lhs = __dmul_rn(q, s);             // 16 * rn(q * (q0 + q))
  double t = __dadd_rn(cx, 2.0);            // 2 * (re + 1)
  double b = __fma_rn(t, t, i2);            // 4 * fma(t, t, i2)
  return (lhs < i2) || (b < 0.25);
}

// Conservative membership test for the period-3 hyperbolic components (the two period-3 bulbs and
// the cardioid of the period-3 copy at c = -1.7549), SURVEY.md 8(f)3.  On those components the
// multiplier lambda of the attracting 3-cycle satisfies
//     c^3 + 2 c^2 + (1 - lambda/8) c + (1 - lambda/8)^2 = 0          (Giarrusso & Fisher 1995),
// a quadratic in mu = 1 - lambda/8, so  lambda = 8 + 4c -+ 4c sqrt(-7 - 4c), and c lies in a
// period-3 component iff one of the two roots has |lambda| < 1.  A candidate with
// |lambda|^2 < kP3Max = 0.998 sits inside: its critical orbit converges to the cycle (every
// attracting cycle attracts the critical orbit) and the cycle stays ~1e-4 away from the Julia set,
// 12 orders of magnitude more than the rounding noise of the FP64 iteration, so the reference's
// loop runs to max_iterations whatever that is.  It is `hit max` without being iterated.  The
// bound is deliberately close to 1: the samples between 0.96 and 0.998 are 1.7 % of the
// never-escaping ones but 12 % of their iterations (12 000 each before they turn bit-periodic).
// FP32 is ample: |lambda|^2 is formed to ~2e-5.  Histogram-neutral like the periodicity check;
// BUDDHA_F_NO_SHORTCUT turns both off and the parity tests run both ways.
constexpr float kP3Max = 0.998f;

__device__ __forceinline__ bool in_period3_component(double cx2, double cy2) {
  const float a = 0.5f * (float)cx2, b = 0.5f * (float)cy2;            // c = a + b i
  const float wr = __fmaf_rn(-4.0f, a, -7.0f), wi = -4.0f * b;          // w = -7 - 4c
  const float mw = sqrtf(__fmaf_rn(wr, wr, wi * wi));
  const float sr = sqrtf(fmaxf(0.5f * (mw + wr), 0.0f));               // s = sqrt(w), principal
  const float si = copysignf(sqrtf(fmaxf(0.5f * (mw - wr), 0.0f)), wi);
  const float tr = __fmaf_rn(a, sr, -b * si), ti = __fmaf_rn(a, si, b * sr);  // t = c s
  const float br = __fmaf_rn(4.0f, a, 8.0f), bi = 4.0f * b;             // 8 + 4c
  const float l1r = __fmaf_rn(-4.0f, tr, br), l1i = __fmaf_rn(-4.0f, ti, bi);
  const float l2r = __fmaf_rn(4.0f, tr, br), l2i = __fmaf_rn(4.0f, ti, bi);
  const float m1 = __fmaf_rn(l1r, l1r, l1i * l1i), m2 = __fmaf_rn(l2r, l2r, l2i * l2i);
  return fminf(m1, m2) < kP3Max;
}

// The same for the six period-4 components.  With mu = lambda / 16 the multipliers of the three
// 4-cycles are the roots of
//     mu^3 - (3 - c^2) mu^2 + (3 + c^2 - c^3 - c^4) mu - (1 + 2c^2 + 3c^3 + 3c^4 + 3c^5 + c^6) = 0
// (coefficients = elementary symmetric functions of the three multipliers, fitted as integer
// polynomials in c from 60-digit cycle computations and checked at further points).  Only a root
// with |lambda| < 0.999 matters: two Newton steps from mu = 0 find it if it exists, and a root of a
// cubic lies within 3 |p / p'| of any point, so |mu| + 3 |p(mu) / p'(mu)| < 0.999 / 16 proves one.
// FP32: the coefficients are < 300 in size, errors ~1e-5 against a margin of 6e-5 in mu.
// It flags 17.5 % of the never-escaping samples (0 escapes among 195 000 flagged in the CPU
// check) and, with the period-3 test, cuts the executed iterations per candidate on config 2
// from 20.2 to 14.4.  ~100 FP32 instructions per `late` batch: worth +3.8 % in the 80-register
// build; in the 72-register build of the tiled contexts and in the fused build it costs more than
// it saves (register pressure), so only render_persistent_kernel<plain or ship, kRegsWide> uses it.
constexpr int kP4NewtonSteps = 2;
constexpr int kPeriodTestMinIt = 512;   // the component tests run only for -m at least this large
struct cfloat { float r, i; };
__device__ __forceinline__ cfloat cmul(cfloat a, cfloat b) {
  return {__fmaf_rn(a.r, b.r, -a.i * b.i), __fmaf_rn(a.r, b.i, a.i * b.r)};
}
__device__ __forceinline__ cfloat cadd(cfloat a, cfloat b) { return {a.r + b.r, a.i + b.i}; }
__device__ __forceinline__ cfloat cdiv(cfloat a, cfloat b) {
  const float d = 1.0f / __fmaf_rn(b.r, b.r, b.i * b.i);
  return {__fmaf_rn(a.r, b.r, a.i * b.i) * d, __fmaf_rn(a.i, b.r, -a.r * b.i) * d};
}
__device__ __forceinline__ bool in_period4_component(double cx2, double cy2) {
  const cfloat c = {0.5f * (float)cx2, 0.5f * (float)cy2};
  const cfloat c2 = cmul(c, c), c3 = cmul(c2, c), c4 = cmul(c2, c2), c5 = cmul(c4, c), c6 = cmul(c3, c3);
  const cfloat a2 = {c2.r - 3.0f, c2.i};
  const cfloat a1 = {3.0f + c2.r - c3.r - c4.r, c2.i - c3.i - c4.i};
  const cfloat a0 = {-(1.0f + 2.0f * c2.r + 3.0f * (c3.r + c4.r + c5.r) + c6.r),
                     -(2.0f * c2.i + 3.0f * (c3.i + c4.i + c5.i) + c6.i)};
  cfloat mu = {0.0f, 0.0f}, p = a0, dp = a1;
#pragma unroll
  for (int k = 0; k < kP4NewtonSteps; k++) {
    const cfloat q = cdiv(p, dp);
    mu = {mu.r - q.r, mu.i - q.i};
    p = cadd(cmul(cadd(cmul(cadd(mu, a2), mu), a1), mu), a0);
    dp = cadd(cmul(cadd({3.0f * mu.r, 3.0f * mu.i}, {2.0f * a2.r, 2.0f * a2.i}), mu), a1);
  }
  const cfloat q = cdiv(p, dp);
  const float bound = sqrtf(__fmaf_rn(mu.r, mu.r, mu.i * mu.i)) + 3.0f * sqrtf(__fmaf_rn(q.r, q.r, q.i * q.i));
  return bound < 0.999f / 16.0f;   // |lambda| < 0.999 (NaN compares false)
}

// ---- scatter --------------------------------------------------------------------------------

// (d) fire-and-forget increment of one cell.  Histograms that fit L2 (or whose hot region does)
// take the reduction directly.  For histograms far beyond L2 (config 3: 1.6 GB) a random 4-byte
// reduction costs a 32-byte sector read-modify-write in HBM (measured ceiling 2.1e10 red/s against
// 1.9e11 red/s inside L2), so the increment is instead APPENDED to a list and applied later by
// apply_tile_kernel tile by tile (64 MB of histogram per tile), while that tile is L2-resident.
// Every (tile, warp) pair owns a private list, so the append needs no global atomic: the slot
// comes from a per-warp counter in shared memory.  (Shared global append counters were measured
// first: one hot address sustains only a few 1e7 atomics-with-return per second, which capped
// the scatter at 4.7e9 .. 1.4e10 points/s, below the direct path.)  A full list falls back to
// the direct reduction, so list capacity only affects speed, never the result.
struct Sink {
  uint32_t *hist;
  uint2 *tile_tab;     // shared memory, one entry per tile: (next pool slot, end of this warp's list)
  uint32_t gwarp;      // global warp index = list column
};

// kDirect: the build of contexts that never tile (the 80-register one): the reduction only.
template <bool kDirect = false>
__device__ __forceinline__ void scatter(const RenderParams &p, const Sink &k, uint32_t idx) {
  if (!kDirect && p.tile_shift) {
    uint2 *e = k.tile_tab + (idx >> p.tile_shift);
    const uint32_t slot = atomicAdd(&e->x, 1u);
    if (slot < e->y) {
      __stcs(p.pool + slot, idx & ((1u << p.tile_shift) - 1u));
      return;
    }
  }
  red_add_u32(k.hist + idx);
}

// Per-warp list table: set it up (mode 0: empty lists, mode 1: continue where an earlier kernel
// of the same launch stopped) and publish the attempted-append counts for apply_tile_kernel and
// the calibration (mode 2).  Not inlined (arguments by value): runs once per warp and launch.
__device__ __noinline__ void tile_table_copy(uint2 *tab, uint32_t *tcount, const uint32_t *tile_cap,
                                             const uint32_t *tile_base, int n_tiles,
                                             uint32_t n_warps, uint32_t gwarp, int mode) {
  __syncwarp();
#pragma unroll 1
  for (int t = lane_id(); t < n_tiles; t += 32) {
    uint32_t *g = tcount + (size_t)t * n_warps + gwarp;
    const uint32_t cap = tile_cap[t];
    const uint32_t begin = tile_base[t] + gwarp * cap;
    if (mode == 2) *g = tab[t].x - begin;
    else tab[t] = make_uint2(begin + (mode == 0 ? 0u : *g), begin + cap);
  }
  __syncwarp();
}

__device__ __forceinline__ void tile_counters_load(const RenderParams &p, const Sink &k, bool zero) {
  if (p.tile_shift)
    tile_table_copy(k.tile_tab, p.tcount, p.tile_cap, p.tile_base, p.n_tiles, p.n_warps, k.gwarp,
                    zero ? 0 : 1);
}

__device__ __forceinline__ void tile_counters_store(const RenderParams &p, const Sink &k) {
  if (p.tile_shift)
    tile_table_copy(k.tile_tab, p.tcount, p.tile_cap, p.tile_base, p.n_tiles, p.n_warps, k.gwarp, 2);
}

// ---- binning --------------------------------------------------------------------------------

// IncrementPixelCounter (cudabrot.cu:302-314) verbatim in arithmetic: IEEE subtract, IEEE divide,
// cvt.rzi (saturating), 32-bit index.
__device__ __forceinline__ bool bin_exact_index(double x2, double y2, const RenderParams &p,
                                                uint32_t *idx) {
  double re = __dmul_rn(x2, 0.5), im = __dmul_rn(y2, 0.5);
  if ((re < p.min_re) || (im < p.min_im)) return false;
  int col = __double2int_rz(__ddiv_rn(__dsub_rn(re, p.min_re), p.delta_re));
  int row = __double2int_rz(__ddiv_rn(__dsub_rn(im, p.min_im), p.delta_im));
  if ((row >= 0) && (row < p.h) && (col >= 0) && (col < p.w)) {
    *idx = (uint32_t)((row * p.w) + col);
    return true;
  }
  return false;
}

template <bool kDirect = false>
__device__ __forceinline__ bool bin_exact(double x2, double y2, const RenderParams &p,
                                          const Sink &hist) {
  uint32_t idx;
  if (!bin_exact_index(x2, y2, p, &idx)) return false;
  scatter<kDirect>(p, hist, idx);
  return true;
}

// ---- the simple kernel (debug / cross-check): one sample per thread, reference dataflow ------

__global__ void __launch_bounds__(256)
render_simple_kernel(RenderParams p, unsigned long long first, uint32_t *__restrict__ hist,
                     unsigned long long *__restrict__ counters) {
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long n_rej = 0, n_hit = 0, n_early = 0, n_acc = 0, e_ref = 0, pts = 0, inc = 0;
  for (unsigned long long s = first + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
       s < p.end; s += stride) {
    uint4 r = philox4x32_10(s, p);
    double cre = __dmul_rn(coord2_from_words(r.x, r.y), 0.5);
    double cim = __dmul_rn(coord2_from_words(r.z, r.w), 0.5);
    // cudabrot.cu:284-298, SASS dataflow of SURVEY.md 8(c)
    double i2 = __dmul_rn(cim, cim);
    double q0 = __dadd_rn(cre, -0.25);
    double q = __fma_rn(q0, q0, i2);
    bool card = __dmul_rn(q, __dadd_rn(q0, q)) < __dmul_rn(i2, 0.25);
    double t = __dadd_rn(cre, 1.0);
    bool bulb = __fma_rn(t, t, i2) < 0.0625;
    if (!p.ship && (card || bulb)) { n_rej++; continue; }
    // cudabrot.cu:319-340
    double re = cre, im = cim;
    int i = p.max_it;
    for (int k = 0; k < p.max_it; k++) {
      if (p.ship) { re = fabs(re); im = fabs(im); }
      double t1 = __dmul_rn(im, im);
      double t2 = __fma_rn(re, re, -t1);
      double r2 = __dadd_rn(re, re);
      im = __fma_rn(r2, im, cim);
      re = __dadd_rn(cre, t2);
      if (__fma_rn(im, im, __dmul_rn(re, re)) > 4.0) { i = k; break; }
    }
    if (i >= p.max_it) { n_hit++; e_ref += (p.max_it > 0 ? p.max_it : 0); continue; }
    e_ref += i + 1;
    if (i < p.min_it) { n_early++; continue; }
    n_acc++;
    // cudabrot.cu:347-365
    re = cre; im = cim;
    for (;;) {
      if (p.ship) { re = fabs(re); im = fabs(im); }
      double t1 = __dmul_rn(im, im);
      double t2 = __fma_rn(re, re, -t1);
      double r2 = __dadd_rn(re, re);
      im = __fma_rn(r2, im, cim);
      re = __dadd_rn(cre, t2);
      pts++;
      if (bin_exact(__dmul_rn(re, 2.0), __dmul_rn(im, 2.0), p, Sink{hist, nullptr, 0u})) inc++;
      if (__fma_rn(im, im, __dmul_rn(re, re)) > 4.0) break;
    }
  }
  unsigned long long v[kCntSlots] = {n_rej, n_hit, n_early, n_acc, e_ref, pts, inc, e_ref, 0, pts};
#pragma unroll
  for (int k = 0; k < kCntSlots; k++) {
    unsigned long long x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    if (lane_id() == 0 && x) atomicAdd(counters + k, x);
  }
}

// ---- the persistent renderer ----------------------------------------------------------------
//
// One warp is the unit of scheduling.  It owns five stacks in shared memory and moves through
// six phases.  The escape test of a fresh candidate is a chain of fixed-length, straight-line
// TIERS: all 32 lanes run the same number of steps with the exact per-step test and no per-lane
// refill, survivors are ballot-compacted onto the next stack.  The survival curve is so flat
// (33 % survive step 1, 19 % step 2, 6.5 % step 6, 2.3 % step 22, 1.5 % step 46) that a tier
// keeps most of its lane-steps useful, while the bookkeeping per step drops to one predicated
// add and one predicate update.
//
//   gen     draw 32 candidates (Philox); FP32 pre-classification retires the certainly rejected
//           and the certain step-1 / step-2 escapes (82 %) without FP64   -> t0 (Philox words)
//   tier 0  c from the words in FP64, exact cardioid/bulb test, steps 1..6 -> t2 (c only)
//   tier 2  steps 7..22 (re-computes steps 1..6)                             -> late
//   late    24 per-step-tested steps from a stored state (c, z, it): tier-2 survivors, samples
//           handed back by deep, tails that have fewer than kBlock steps left -> deep / late
//   deep    long escape tests: kBlock unchecked steps (4 FP64 instr each) per round, one |z|^2
//           test per round, exact periodicity shortcut, per-lane refill.  A lane whose round
//           ended outside the radius-2 disc is handed to `late` with its round-start state.
//   orbit   re-iterate accepted samples for exactly i+1 steps and scatter with red.global.add.
//
// Every tier may also push accepted escapes to `orbit`.  The scheduler serves the stacks in the
// order orbit, late, deep, t2, t0, gen; a phase runs only while each stack it pushes to holds
// < 32 entries (late re-queues its own survivors while deep is full), so no stack exceeds 63
// entries and some phase can always run.
//
// Lanes that have escaped keep stepping until their tier ends; their values grow to inf/NaN,
// which no later code reads (the `alive` predicate is sticky and NaN compares false).

// Shared-memory work stacks of one warp (structure of arrays: 16-byte accesses, no conflicts).
//   c_*      t0 (slots 0 up): Philox words of candidates the sampler's float pre-test could not
//            retire; t2 (slots kCCap-1 down): candidates that re-compute their state from c
//   z_*      late (0 up): it = iterations done; orb (kZCap-1 down): it = steps still to record
//   deep_*   carry their checkpoint, so suspending a lane does not restart the periodicity search
//            (long cycles need long uninterrupted windows); meta = (last, age): the age after which
//            no further full round fits below max_it, and the rounds spent in deep so far
struct WarpQueues {
  double2 deep_c[kDeepCap], deep_z[kDeepCap], deep_r[kDeepCap];
  double2 z_c[kZCap], z_z[kZCap];
  union { double2 c_c[kCCap]; uint4 c_w[kCCap]; };
  uint2 deep_meta[kDeepCap];
  int z_it[kZCap];
  int pad[(4 - (2 * kDeepCap + kZCap) % 4) % 4];  // keeps the next warp's arrays 16-byte aligned
  // fused render: this warp's per-channel / per-band accumulators (ChannelSlot), added to the
  // global ones when the kernel ends.  (They used to be global atomics from every late / tier
  // batch: 6 same-address atomics per ~1000 candidates from 10 000 warps.)
  unsigned long long ch_cnt[kChCount];
};
static_assert(sizeof(WarpQueues) % 16 == 0, "stack arrays must stay 16-byte aligned");

// slot of entry i of the stack growing up from 0 / down from the top of a shared array
template <bool kDown, int kCap>
__device__ __forceinline__ int slot_of(int i) { return kDown ? kCap - 1 - i : i; }
constexpr bool kLate = false, kOrb = true;  // the two stacks in z_*
constexpr bool kT0 = false, kT2 = true;     // the two stacks in c_*

// Per-warp state that lives in registers for the whole kernel.
struct WarpState {
  int t0_n, t2_n, late_n, deep_n, orb_n;  // stack heights (warp-uniform)
  int pend_n, fail_n;                      // the warp's certificate queue in global memory (see cert_phase)
  unsigned long long chunk_base;           // first sample index of the chunk this warp owns
  uint32_t chunk_off, chunk_len;           // progress inside the chunk
  bool exhausted;                          // the global cursor ran past p.end
  // per-lane event counters, flushed to global memory at every cursor grab
  uint32_t n_rej, n_hit, n_acc, n_cyc, n_exact;
  uint32_t steps;      // iterations that advanced a sample (count for escape_iters AND executed)
  uint32_t skipped;    // iterations the periodicity shortcut did not have to run (escape_iters only)
  // deep rounds: a lane subtracts a sample's age when it loads it and adds the age back when the
  // sample finishes or is suspended, so the sum is the number of rounds run (transiently negative
  // per lane, hence signed); d_out = rounds rolled back because the sample had escaped in them.
  // escape_iters += kBlock * (d_rounds - d_out), executed += kBlock * d_rounds.
  int32_t d_rounds;
  uint32_t d_out;
  uint32_t p_pts, p_inc;
  uint32_t ch_inc[kMaxBands];     // fused render only: increments per band
};

// Warp-reduce the per-lane counters and add them to the global accumulators.  Deliberately not
// inlined (arguments by value, so the caller's state stays in registers): it runs a few times per
// launch and would otherwise be replicated at every call site.
__device__ __noinline__ void flush_values(unsigned long long *counters, uint32_t n_rej,
                                          uint32_t n_hit, uint32_t n_acc, uint32_t steps,
                                          uint32_t skipped, uint32_t p_pts, uint32_t p_inc,
                                          int32_t d_rounds, uint32_t d_out, uint32_t n_cyc,
                                          uint32_t n_exact) {
  // (64-bit two's complement: a transiently negative d_rounds adds up correctly)
  const long long deep_exec = (long long)d_rounds * kBlock;
  const long long deep_ref = deep_exec - (long long)d_out * kBlock;
  const unsigned long long v[kCntSlots] = {
      n_rej, n_hit, 0ull, n_acc,
      (unsigned long long)steps + (unsigned long long)skipped + (unsigned long long)deep_ref,
      p_pts, p_inc, (unsigned long long)steps + (unsigned long long)deep_exec, n_cyc, n_exact};
#pragma unroll 1
  for (int k = 0; k < kCntSlots; k++) {
    if (k == kCntTooEarly) continue;  // derived on the host: candidates - the other classes
    unsigned long long x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    if (lane_id() == 0 && x) atomicAdd(counters + k, x);
  }
}

__device__ __forceinline__ void flush_counters(WarpState &ws, unsigned long long *counters) {
  flush_values(counters, ws.n_rej, ws.n_hit, ws.n_acc, ws.steps, ws.skipped, ws.p_pts, ws.p_inc,
               ws.d_rounds, ws.d_out, ws.n_cyc, ws.n_exact);
  ws.n_rej = ws.n_hit = ws.n_acc = ws.n_cyc = ws.n_exact = 0;
  ws.steps = ws.skipped = ws.d_out = ws.p_pts = ws.p_inc = 0;
  ws.d_rounds = 0;
}

// Adds v (summed over the warp) to one per-channel accumulator; called from rare branches only.
// chc = the warp's accumulators in shared memory (render kernel) or the global ones (drain).
__device__ __forceinline__ void channel_add(unsigned long long *chc, int ch, int slot, uint32_t v) {
  v = __reduce_add_sync(kFull, v);
  if (lane_id() == 0 && v) atomicAdd(chc + ch * kChSlots + slot, (unsigned long long)v);
}

// Fused render: the per-channel increment counters.
template <int kVar>
__device__ __forceinline__ void flush_channel_counters(const RenderParams &p, WarpState &ws,
                                                       unsigned long long *chc) {
  if constexpr ((kVar & kVarFused) != 0) {
#pragma unroll
    for (int k = 0; k < kMaxBands; k++) {
      if (k < p.n_bands) channel_add(chc, k, kChIncrements, ws.ch_inc[k]);
      ws.ch_inc[k] = 0;
    }
  }
}

// Ballot-compacted push of one entry per lane with `pred` set.
__device__ __forceinline__ int push_slot(int &height, bool pred) {
  unsigned m = __ballot_sync(kFull, pred);
  int slot = height + __popc(m & lanemask_lt());
  height += __popc(m);
  return slot;
}

template <bool kDown>
__device__ __forceinline__ void push_z(WarpQueues &q, int &height, bool pred, double cx, double cy,
                                       double x, double y, int it) {
  const int slot = slot_of<kDown, kZCap>(push_slot(height, pred));
  if (pred) { q.z_c[slot] = make_double2(cx, cy); q.z_z[slot] = make_double2(x, y); q.z_it[slot] = it; }
}

// Band (+1; 0 = no channel) that an escape at step `it` (1-based count, it <= max_it) falls into.
__device__ __forceinline__ unsigned accept_band(const RenderParams &p, int it) {
  unsigned code = 0;
#pragma unroll
  for (int s = 0; s < 2 * kMaxChannels; s++)
    if (s < p.n_seg && it >= p.seg_start[s] && it < p.seg_start[s + 1])
      code = (unsigned)(p.seg_band[s] + 1);
  return code;
}

// The accept filter (cudabrot.cu:407-408) for lanes whose sample escaped at step `it`, and the
// push of the accepted ones.  Fused render: a sample is pushed if any channel accepts it; its
// band travels in the top bits of the step count.
template <int kVar>
__device__ __forceinline__ void push_orbit(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                           unsigned long long *counters, bool esc, double cx,
                                           double cy, int it) {
  unsigned code = 0;  // band + 1
  bool acc;
  if constexpr ((kVar & kVarFused) != 0) {
    code = esc ? accept_band(p, it) : 0u;
    acc = code != 0u;
  } else {
    acc = esc && it - 1 >= p.min_it;
  }
  if (__ballot_sync(kFull, acc) == 0u) return;
  ws.n_acc += acc ? 1u : 0u;
  ws.p_pts += acc ? (uint32_t)it : 0u;
  if (__ballot_sync(kFull, (ws.p_pts >> 30) != 0u)) flush_counters(ws, counters);  // (huge -m)
  int n = it;
  if constexpr ((kVar & kVarFused) != 0) {
    for (int k = 0; k < p.n_bands; k++) {
      const bool a = code == (unsigned)(k + 1);
      const unsigned n = __popc(__ballot_sync(kFull, a));
      if (n == 0u) continue;
      if (lane_id() == 0) q.ch_cnt[k * kChSlots + kChAccepted] += n;
      channel_add(q.ch_cnt, k, kChPoints, a ? (uint32_t)it : 0u);
    }
    n |= (int)((code - 1u) << kOrbStepBits);
  }
  push_z<kOrb>(q, ws.orb_n, acc, cx, cy, cx, cy, n);
}

// Fused render: a sample ESCAPED at step it_f.  Channels whose limit lies below it_f count it as
// "hit max" and must not be charged the iterations beyond their limit (escape_iters_k = sum of
// min(it_f, ch_max[k])).  Samples that never escape hit every channel's limit; the host adds
// those from the common hit counter (buddha_get_channel_counters), so deep stays untouched.
template <int kVar>
__device__ __forceinline__ void channel_finish(const RenderParams &p, WarpQueues &q, bool esc,
                                               int it_f) {
  if constexpr ((kVar & kVarFused) != 0) {
    const bool any = esc && it_f > p.ch_low;
    if (__ballot_sync(kFull, any) == 0u) return;
    for (int k = 0; k < p.n_ch; k++) {
      const bool over = any && it_f > p.ch_max[k];
      const unsigned n = __popc(__ballot_sync(kFull, over));
      if (n == 0u) continue;
      if (lane_id() == 0) q.ch_cnt[k * kChSlots + kChHit] += n;
      channel_add(q.ch_cnt, k, kChOver, over ? (uint32_t)(it_f - p.ch_max[k]) : 0u);
    }
  }
}

// N steps with the exact per-step escape test (cudabrot.cu:331-338) for all lanes at once.
// *cnt = steps a lane ran while it had not escaped (the escaping step included); `alive` on
// return: never escaped.  A step limit below N is applied by the caller afterwards.
template <int kVar, int N>
__device__ __forceinline__ void tested_steps(double &x, double &y, double cx, double cy,
                                             bool &alive, int &cnt) {
  // Eight steps per loop trip where N allows (tier 2: 16, late: 24): 200 instructions less hot code
  // for one loop branch per eight steps (config 1 +0.6 %, config 4 +0.8 %, config 2 +-0;
  // profiles/r02_tested_unroll_ab.txt).  BUDDHA_TESTED_UNROLL=0: fully unrolled (A/B builds).
#ifndef BUDDHA_TESTED_UNROLL
#define BUDDHA_TESTED_UNROLL 8
#endif
  constexpr int kU = (BUDDHA_TESTED_UNROLL > 0 && N % BUDDHA_TESTED_UNROLL == 0) ? BUDDHA_TESTED_UNROLL : N;
#pragma unroll 1
  for (int k0 = 0; k0 < N; k0 += kU) {
#pragma unroll
    for (int k = 0; k < kU; k++) {
      BUDDHA_ZSTEP(x, y, cx, cy);
      const bool e = norm4(x, y) > 16.0;
      cnt += alive ? 1 : 0;
      alive = alive && !e;
    }
  }
}

// (a) sampler.  One batch = one candidate per lane: Philox, then an FP32 PRE-CLASSIFICATION that
// retires 82 % of the candidates without a single FP64 instruction (which cost two issue cycles
// each on this machine, profiles/r02_issue_probes.txt): coordinates from the top 23 bits of the two
// high Philox words, the cardioid / bulb test and the first two steps in float.  A decision counts
// only when it clears its threshold by a margin 50 x the largest float error seen over 2^26
// samples (CPU check in the test suite: no decided sample disagrees with the reference FP64 arithmetic):
//   certainly rejected                      -> counted, done
//   certainly outside radius 2 after step 1 -> escaped at step 1 (an escaping c is far from the
//   certainly inside after step 1, outside     cardioid and the bulb, so it was not rejected): too
//   after step 2                               early for any cutoff >= 2, counted, done
// Everything else -- samples still inside after two steps and the few undecided ones -- goes to
// `t0` as its four Philox words; the first exact tier re-derives c from them in FP64.
constexpr float kPreRejMargin = 0.02f, kPreEscMargin = 0.05f;

// The sampler's FP32 pre-classification of one candidate from its Philox words (see gen_phase).
template <int kVar>
__device__ __forceinline__ void gen_classify(const uint4 &r, float esc1_min, float esc2_min,
                                             bool &rej, bool &esc1, bool &esc2) {
  // 2c to ~1e-6: the mantissa trick turns the top 23 bits of a word into [1, 2)
  const float cx = __fmaf_rn(__uint_as_float(0x3F800000u | (r.y >> 9)), 8.0f, -12.0f);
  const float cy = __fmaf_rn(__uint_as_float(0x3F800000u | (r.w >> 9)), 8.0f, -12.0f);
  const float i2 = cy * cy;
  rej = false;
  if constexpr ((kVar & kVarShip) == 0) {  // rejected2 in float (cudabrot.cu:284-298, :397-399)
    const float q0 = cx - 0.5f, qq = __fmaf_rn(q0, q0, i2), sx = __fmaf_rn(q0, 2.0f, qq);
    const float t = cx + 2.0f;
    rej = (qq * sx < i2 - kPreRejMargin) || (__fmaf_rn(t, t, i2) < 0.25f - kPreRejMargin);
  }
  const float x1 = __fmaf_rn(__fmaf_rn(cx, cx, -i2), 0.5f, cx);
  const float y1 = (kVar & kVarShip) ? __fmaf_rn(fabsf(cx), fabsf(cy), cy) : __fmaf_rn(cx, cy, cy);
  const float n1 = __fmaf_rn(y1, y1, x1 * x1);
  const float x2 = __fmaf_rn(__fmaf_rn(x1, x1, -y1 * y1), 0.5f, cx);
  const float y2 = (kVar & kVarShip) ? __fmaf_rn(fabsf(x1), fabsf(y1), cy) : __fmaf_rn(x1, y1, cy);
  const float n2 = __fmaf_rn(y2, y2, x2 * x2);
  esc1 = !rej && n1 > esc1_min;
  esc2 = !rej && n1 < 16.0f - kPreEscMargin && n2 > esc2_min;
}

template <int kVar>
__device__ __forceinline__ void gen_phase(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                          unsigned long long *cursor,
                                          unsigned long long *counters) {
  const unsigned lane = lane_id();
  // an escape at step k is "too early" (and inside the limit) for every sample only if
  // min_it >= k and max_it >= k; otherwise the exact tier has to look at it
  const bool quick1 = p.min_it >= 1 && p.max_it >= 1, quick2 = p.min_it >= 2 && p.max_it >= 2;
  // (folded into the thresholds: nothing exceeds +inf, and no predicate has to be kept per batch)
  const float esc1_min = quick1 ? 16.0f + kPreEscMargin : __int_as_float(0x7f800000);
  const float esc2_min = quick2 ? 16.0f + kPreEscMargin : __int_as_float(0x7f800000);
#pragma unroll 1
  while (ws.t0_n < 32) {
    if (ws.chunk_off >= ws.chunk_len) {
      if (ws.exhausted) break;
      // The per-lane counters go to global memory when one of them nears 2^30 (and at the end
      // of the kernel), not at every grab: 10 same-address atomics per warp and grab were
      // measurable once the L2 is busy with reductions that miss (10000x10000: +25..70 %).
      {
        uint32_t any = ws.n_rej | ws.n_hit | ws.n_acc | ws.n_cyc | ws.n_exact | ws.steps |
                       ws.skipped | ws.d_out | ws.p_pts | ws.p_inc | (uint32_t)abs(ws.d_rounds);
#pragma unroll
        for (int k = 0; k < kMaxBands; k++) any |= ws.ch_inc[k];
        if (__ballot_sync(kFull, (any >> 29) != 0u)) {
          flush_counters(ws, counters);
          flush_channel_counters<kVar>(p, ws, q.ch_cnt);
        }
      }
      unsigned long long base = 0;
      if (lane == 0) base = atomicAdd(cursor, (unsigned long long)p.chunk);
      base = __shfl_sync(kFull, base, 0);
      if (base >= p.end) { ws.exhausted = true; break; }
      ws.chunk_base = base;
      ws.chunk_off = 0;
      ws.chunk_len = (base + p.chunk <= p.end) ? p.chunk : (uint32_t)(p.end - base);
    }
    if constexpr (kGen2) {
      // two independent candidates per lane: twice the instruction-level parallelism in the Philox
      // chains, one loop trip and one stack update for both
      const uint32_t o0 = ws.chunk_off + lane, o1 = o0 + 32u;
      const bool valid0 = o0 < ws.chunk_len, valid1 = o1 < ws.chunk_len;
      ws.chunk_off += 64;
      const uint4 r0 = philox4x32_10(ws.chunk_base + o0, p);
      const uint4 r1 = philox4x32_10(ws.chunk_base + o1, p);
      bool rej0, e10, e20, rej1, e11, e21;
      gen_classify<kVar>(r0, esc1_min, esc2_min, rej0, e10, e20);
      gen_classify<kVar>(r1, esc1_min, esc2_min, rej1, e11, e21);
      ws.n_rej += ((valid0 && rej0) ? 1u : 0u) + ((valid1 && rej1) ? 1u : 0u);
      ws.steps += ((valid0 && e10) ? 1u : 0u) + ((valid1 && e11) ? 1u : 0u);
      ws.steps += ((valid0 && e20) ? 2u : 0u) + ((valid1 && e21) ? 2u : 0u);
      const bool keep0 = valid0 && !(rej0 || e10 || e20), keep1 = valid1 && !(rej1 || e11 || e21);
      const unsigned m0 = __ballot_sync(kFull, keep0), m1 = __ballot_sync(kFull, keep1);
      const int s0 = ws.t0_n + __popc(m0 & lanemask_lt());
      const int s1 = ws.t0_n + __popc(m0) + __popc(m1 & lanemask_lt());
      ws.t0_n += __popc(m0) + __popc(m1);
      if (keep0) q.c_w[slot_of<kT0, kCCap>(s0)] = r0;
      if (keep1) q.c_w[slot_of<kT0, kCCap>(s1)] = r1;
    } else {
    const uint32_t o = ws.chunk_off + lane;
    const bool valid = o < ws.chunk_len;
    ws.chunk_off += 32;
    const uint4 r = philox4x32_10(ws.chunk_base + o, p);
    bool rej, esc1, esc2;
    gen_classify<kVar>(r, esc1_min, esc2_min, rej, esc1, esc2);
    ws.n_rej += (valid && rej) ? 1u : 0u;
    ws.steps += (valid && esc1) ? 1u : 0u;
    ws.steps += (valid && esc2) ? 2u : 0u;
    const bool keep = valid && !(rej || esc1 || esc2);
    const int slot = slot_of<kT0, kCCap>(push_slot(ws.t0_n, keep));
    if (keep) q.c_w[slot] = r;
    }
  }
  __syncwarp();
}

// (b0) the first exact tier: pops up to 32 candidates the sampler could not retire, derives c from
// their Philox words in FP64, applies the exact cardioid / bulb test and runs steps 1..kT1End with
// the exact per-step test.  Survivors go to t2 as c only.
template <int kVar>
__device__ __forceinline__ void first_tier_phase(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                                 unsigned long long *counters) {
  const int take = min(ws.t0_n, 32);
  const bool act = (int)lane_id() < take;
  ws.t0_n -= take;
  uint4 r = make_uint4(0u, 0u, 0u, 0u);
  if (act) r = q.c_w[slot_of<kT0, kCCap>(ws.t0_n + (int)lane_id())];
  const double cx = coord2_from_words(r.x, r.y);
  const double cy = coord2_from_words(r.z, r.w);
  const bool rej = (kVar & kVarShip) ? false : rejected2(cx, cy);  // cudabrot.cu:397-399
  ws.n_rej += (act && rej) ? 1u : 0u;
  const bool cand = act && !rej;
  double x = cx, y = cy;
  bool alive = cand;
  int cnt = 0;
  tested_steps<kVar, kT1End>(x, y, cx, cy, alive, cnt);
  if (p.max_it > kT1End) {
    // the common case: every step counts, survivors move on
    ws.steps += (uint32_t)cnt;
    const int slot = slot_of<kT2, kCCap>(push_slot(ws.t2_n, alive));
    if (alive) q.c_c[slot] = make_double2(cx, cy);
    if (kT1End - 1 >= p.min_it) push_orbit<kVar>(p, q, ws, counters, cand && !alive, cx, cy, cnt);
  } else {
    // max_it falls inside this tier (or is 0): escapes after the limit do not count, the rest
    // has hit max (IterateMandelbrot stops at max, cudabrot.cu:326)
    const int allowed = max(p.max_it, 0);
    const bool esc = cand && !alive && cnt <= allowed;
    ws.steps += cand ? (uint32_t)min(cnt, allowed) : 0u;
    ws.n_hit += (cand && !esc) ? 1u : 0u;
    if (kT1End - 1 >= p.min_it) push_orbit<kVar>(p, q, ws, counters, esc, cx, cy, cnt);
  }
  __syncwarp();
}

__device__ __forceinline__ void push_deep(WarpQueues &q, WarpState &ws, bool pred, double cx,
                                          double cy, double x, double y, double rx, double ry,
                                          unsigned last, unsigned age) {
  unsigned m = __ballot_sync(kFull, pred);
  if (m == 0u) return;
  int slot = ws.deep_n + __popc(m & lanemask_lt());
  ws.deep_n += __popc(m);
  if (pred) {
    q.deep_c[slot] = make_double2(cx, cy); q.deep_z[slot] = make_double2(x, y);
    q.deep_r[slot] = make_double2(rx, ry); q.deep_meta[slot] = make_uint2(last, age);
  }
}

// (b) one tier of the escape test: pops up to 32 candidates that survived A steps, re-computes
// those steps from c (no tests needed: they are known not to escape), runs steps A+1..A+N with the
// per-step test.  kToLate = false: survivors go to t2 as c only; true: to `late` with their state.
template <int kVar, int A, int N, bool kToLate>
__device__ __forceinline__ void tier_phase(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                           int &src_n, unsigned long long *counters) {
  const int take = min(src_n, 32);
  const bool act = (int)lane_id() < take;
  src_n -= take;
  double2 c = make_double2(0.0, 0.0);
  if (act) c = q.c_c[slot_of<kT2, kCCap>(src_n + (int)lane_id())];
  const double cx = c.x, cy = c.y;
  double x = cx, y = cy;
#pragma unroll
  for (int k = 0; k < A; k++) BUDDHA_ZSTEP(x, y, cx, cy);
  bool alive = act;
  int cnt = 0;
  tested_steps<kVar, N>(x, y, cx, cy, alive, cnt);
  if (p.max_it > A + N) {
    // the common case: every step counts, survivors move on
    ws.steps += (uint32_t)cnt;
    if (kToLate) {
      push_z<kLate>(q, ws.late_n, alive, cx, cy, x, y, A + N);
    } else {
      const int slot = slot_of<kT2, kCCap>(push_slot(ws.t2_n, alive));
      if (alive) q.c_c[slot] = make_double2(cx, cy);
    }
    if (A + N - 1 >= p.min_it) push_orbit<kVar>(p, q, ws, counters, act && !alive, cx, cy, A + cnt);
  } else {
    // max_it falls inside this tier: escapes after the limit do not count, the rest has hit max
    const int allowed = p.max_it - A;  // >= 1: nothing is pushed here once max is reached
    const bool esc = act && !alive && cnt <= allowed;
    ws.steps += (uint32_t)min(cnt, allowed);
    ws.n_hit += (act && !esc) ? 1u : 0u;
    if (A + N - 1 >= p.min_it) push_orbit<kVar>(p, q, ws, counters, esc, cx, cy, A + cnt);
  }
  __syncwarp();
}

// (b') kLateSteps per-step-tested steps from a stored state.  Entries: tier-2 survivors (it = 22), samples
// handed back by deep (certain to escape within kBlock steps), tails (fewer than kBlock steps left
// below max), samples that deep cannot take (|c| too close to 2, or deep full right now).
template <int kVar, bool kPeriod4>
__device__ __forceinline__ void late_phase(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                           unsigned long long *counters) {
  const int take = min(ws.late_n, 32);
  const bool act = (int)lane_id() < take;
  ws.late_n -= take;
  double cx = 0.0, cy = 0.0, x = 0.0, y = 0.0;
  int it = 0;
  if (act) {
    const int slot = slot_of<kLate, kZCap>(ws.late_n + (int)lane_id());
    double2 c = q.z_c[slot], z = q.z_z[slot];
    cx = c.x; cy = c.y; x = z.x; y = z.y; it = q.z_it[slot];
  }
  __syncwarp();  // the slots are re-used by the pushes below
  const bool deep_room = ws.deep_n < 32;
  bool alive = act;
  int cnt = 0;
  tested_steps<kVar, kLateSteps>(x, y, cx, cy, alive, cnt);
  const int allowed = min(kLateSteps, p.max_it - it);  // >= 1 for every stored entry
  const bool esc = act && !alive && cnt <= allowed;
  ws.steps += act ? (uint32_t)min(cnt, allowed) : 0u;
  const bool surv = act && !esc;
  const int nit = it + kLateSteps;
  const bool hit = surv && nit >= p.max_it;  // ran all max iterations (allowed == max - it)
  ws.n_hit += hit ? 1u : 0u;
  push_orbit<kVar>(p, q, ws, counters, esc, cx, cy, it + cnt);
  channel_finish<kVar>(p, q, esc, it + cnt);
  const bool cont = surv && !hit;
  if (__ballot_sync(kFull, cont)) {
    // samples inside a period-3 (or period-4) component never escape: hit max without iterating
    bool cont2 = cont;
    // (below ~500 iterations the tests cost more than the iterations they save: config 1 -1.4 %)
    if ((kVar & kVarShip) == 0 && p.shortcut && p.max_it >= kPeriodTestMinIt) {
      bool p3 = cont && in_period3_component(cx, cy);
      if constexpr (kPeriod4) p3 = p3 || (cont && in_period4_component(cx, cy));
      ws.n_hit += p3 ? 1u : 0u;
      ws.n_cyc += p3 ? 1u : 0u;
      ws.skipped += p3 ? (uint32_t)(p.max_it - nit) : 0u;
      cont2 = cont && !p3;
    }
    // deep's no-re-entry argument needs |c| <= 1.99937, and a full unchecked round must fit
    const bool todeep = cont2 && deep_room && norm4(cx, cy) <= 15.99 && nit + kBlock <= p.max_it;
    // meta = (last, age): after `last` rounds fewer than kBlock steps are left below max_it
    push_deep(q, ws, todeep, cx, cy, x, y, x, y, (unsigned)(p.max_it - nit) / kBlock, 0u);
    push_z<kLate>(q, ws.late_n, cont2 && !todeep, cx, cy, x, y, nit);
    if (__ballot_sync(kFull, (ws.skipped >> 30) != 0u)) flush_counters(ws, counters);  // (huge -m)
  }
  __syncwarp();
}

// ---- attracting-cycle certificate ---------------------------------------------------------------
//
// After the closed-form tests for periods 3 and 4 the never-escaping samples that are left (0.33 %
// of the candidates on config 2) still cost 10.5 of the 14.4 iterations a candidate executes: they
// run until their state repeats bit for bit, 3160 iterations on average.  A sample that has spent
// 64, 256, 1024, ... rounds in `deep` is therefore parked for cert_phase, which looks for the
// attracting cycle itself:
//   1. period: the first k <= pmax with |z_k - z| < kCertTol, iterating in float from the sample's
//      current state z;
//   2. Newton on f^p(w) - w in double from w = z, the derivative d = prod 2 w_j carried along the
//      p steps (on the scaled state 2 w = X + iY, so d <- (X + iY) d: four FP64 instructions);
//   3. accepted if the residual |f^p(w) - w| is below kCertResMax and |d|^2 < kCertLam2Max.
// Then f_c has an attracting cycle.  A quadratic polynomial has at most one attracting cycle and it
// attracts the critical orbit (Fatou), so c lies in a hyperbolic component of the Mandelbrot set,
// the orbit of 0 converges to the cycle, and the reference's loop runs to max_iterations whatever
// that is: `hit max` without iterating further.  Histogram-neutral like the periodicity check and
// the period-3/4 tests, and evidence of the same kind: the test suite restates it on the CPU
// (check_certificate) and runs the reference's own loop on every certified sample (0 escapes among
// 161 000 certified in 2 x 2^26 samples from age 64, 0 among 425 000 from age 4; with the
// multiplier bound at 1.1 the same check finds 44 in 2^22), and every GPU parity test runs with it
// (and without: BUDDHA_F_NO_SHORTCUT).  Samples the certificate cannot settle (period > pmax,
// multiplier too close to 1, Newton not converged) go back to `deep` unchanged.
//
// When to try.  From age 4 the certificate settles 96 % of these samples (588 instead of 3160
// iterations each, executed iterations per candidate 14.4 -> 6.1) but config 2 gains only 2 %: a
// saved deep iteration is worth ~0.18 issue cycles per candidate -- the FP64 chains of `deep` run
// largely in issue slots the other warps leave free -- while an attempt costs 150-200, and at age
// 4 four in ten attempts are samples that escape later.  Measured on config 2 at 2^35-sample
// steps: first age 4 +2.2 %, 16 +4.1 %, 64 +6.5 % (36 % settled, 14.4 -> 9.2 iterations), 256
// +1.6 %.  At -m 5000 (config 4) and in the 72-register and fused builds it costs 2-10 %, so only
// render_persistent_kernel<plain, kRegsWide> uses it, for -m >= kCertMinIt.
#ifndef BUDDHA_CERT_FIRST_AGE
#define BUDDHA_CERT_FIRST_AGE 64
#endif
constexpr unsigned kCertFirstAge = BUDDHA_CERT_FIRST_AGE;  // tried at this age (a power of four) and at 4 x, 16 x, ... it
#ifndef BUDDHA_CERT_PMAX_FIRST
#define BUDDHA_CERT_PMAX_FIRST 32
#endif
constexpr int kCertPmaxFirst = BUDDHA_CERT_PMAX_FIRST, kCertPmaxLater = 64;
constexpr int kCertPasses = 5;                   // Newton evaluations (the last one only verifies)
constexpr float kCertTol = 1e-2f;
constexpr double kCertResMax = 1e-12, kCertLam2Max = 0.998;
#ifndef BUDDHA_CERT_MIN_IT
#define BUDDHA_CERT_MIN_IT 10000
#endif
constexpr int kCertMinIt = BUDDHA_CERT_MIN_IT;   // below this -m the bit-exact search is cheaper (-m 5000: -2.4 %, -m 20000: +6.5 %: break-even near 9000)
constexpr int kCertQueue = 512;                  // queue entries per warp (host side: RenderParams::cert_cap)
constexpr int kCertHeadroom = 32;                // entries kept free at the top (cert_phase, stage C)
#ifndef BUDDHA_CERT_GANG
#define BUDDHA_CERT_GANG 12
#endif
constexpr int kCertGang = BUDDHA_CERT_GANG;      // lanes that wait at the end of a Newton pass before it is handled

// the next age at which a sample of age `age` is due for the certificate: the smallest power of
// four >= kCertFirstAge that is > age
__device__ __forceinline__ unsigned next_cert_age(unsigned age) {
  return max(kCertFirstAge, 1u << (((31 - __clz((int)(age | 1u))) & ~1) + 2));
}

// Checkpoint schedule of the periodicity search: a new checkpoint after 1, 2, 3, 4, 6, 8, 12, 16,
// 24, ... rounds (ages with at most two significant bits).  A cycle is found once a checkpoint
// lies on it and the gap to the next checkpoint covers its period (in rounds); the 1.33..1.5
// spacing costs 10 % fewer iterations on in-set samples than Brent's powers of two (simulated on
// the config-2 sample distribution: 2054 vs 2280 iterations per in-set sample, ideal 1795).
__device__ __forceinline__ bool checkpoint_age(unsigned age) {
  const unsigned rest = age & (age - 1u);     // age without its lowest set bit
  return rest == 0u || 3u * rest == 2u * age;  // nothing left, or the lowest bit sits right below
}

// (c) long escape tests.  Each round runs kBlock unchecked steps and tests |z|^2 once.  Because
// |c| <= 2 here, an orbit that leaves the radius-2 disc cannot re-enter it,
// so "escaped somewhere in the round" <=> "outside at the end of the round".  A state that
// repeats bit-for-bit proves the orbit periodic, i.e. it never escapes (exact shortcut).
//
// Every sample in `late` and `deep` has done kT2End + j * kBlock iterations (tier 2 ends at
// kT2End; late batches and deep rounds are kBlock steps long), so a deep entry only carries
// (last, age): it has done m24 - (last - age) * kBlock iterations, m24 = the largest such count
// <= max_it.  The handling of finished lanes runs after ~45 % of the rounds (32 lanes, ~50 rounds
// per sample), so it is kept short: no per-lane iteration counters (see WarpState::d_rounds), and
// finished lanes keep iterating on stale values, which nothing reads (`act` guards every use; a
// stale orbit may run to inf/NaN, which costs nothing on this hardware).
template <int kVar, bool kCert>
__device__ __forceinline__ void deep_phase(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                           bool drain, bool park, unsigned long long *counters,
                                           uint32_t gwarp) {
  static_assert(kLateSteps == kBlock, "late batches and deep rounds must have the same length");
  const int max_it = p.max_it;
  const int m24 = max_it - (max_it - kT2End) % kBlock;
  const bool shortcut = p.shortcut != 0;
  // (park = false once the sample range is used up: a sample parked then would only come back
  //  after everything else has drained, and the tail of a launch is as long as its longest chain)
  const bool cert_on = kCert && park && shortcut && p.cert_cap != 0u;
  bool act = false;
  double cx = 0.0, cy = 0.0, x = 0.0, y = 0.0, rx = 0.0, ry = 0.0;
  unsigned age = 0;     // rounds this sample has spent in deep (checkpoint schedule)
  unsigned last = 0;    // the age after which no further full round fits below max_it
  unsigned ev = 0;      // the next age at which the lane needs attention: min(last, certificate age)
#pragma unroll 1
  for (;;) {
    // keep room for 32 hand-backs
    if (ws.late_n >= 32 || ws.late_n + ws.orb_n > kZJoint) break;
    const unsigned idle = __ballot_sync(kFull, !act);
    if (ws.deep_n > 0 && idle != 0u) {
      const int rank = __popc(idle & lanemask_lt());
      if (!act && rank < ws.deep_n) {
        const int slot = ws.deep_n - 1 - rank;
        const double2 c = q.deep_c[slot], z = q.deep_z[slot], r = q.deep_r[slot];
        const uint2 meta = q.deep_meta[slot];
        cx = c.x; cy = c.y; x = z.x; y = z.y; rx = r.x; ry = r.y;
        last = meta.x; age = meta.y;  // last >= age + 1
        ev = cert_on ? min(last, next_cert_age(age)) : last;
        ws.d_rounds -= (int32_t)age;
        act = true;
      }
      ws.deep_n -= min(__popc(idle), ws.deep_n);
      __syncwarp();
    }
    const unsigned am = __ballot_sync(kFull, act);
    if (am == 0u) break;
    if (!drain && __popc(am) < kDeepExit) break;

    // rounds without bookkeeping until some lane needs attention
    double x0, y0;
    bool out, same_x;
#pragma unroll 1
    do {
      // the checkpoint a lane is due after `age` rounds, taken before the next round (also the
      // one still pending when a lane was suspended)
      if (checkpoint_age(age)) { rx = x; ry = y; }
      x0 = x; y0 = y;
#pragma unroll
      for (int k = 0; k < kBlock; k++) BUDDHA_ZSTEP(x, y, cx, cy);
      out = !(norm4(x, y) <= 16.0);  // also true for NaN / inf
      age++;
      same_x = shortcut && __double_as_longlong(x) == __double_as_longlong(rx);
    } while (__ballot_sync(kFull, act && (out || same_x || age == ev)) == 0u);
    {
      // (the copy hides `age` from the compiler's induction-variable pass, which otherwise keeps
      // `it` and last - age up to date inside the round loop: eight extra adds per round)
      unsigned age_now = age;
      asm volatile("" : "+r"(age_now));
      const bool cyc = same_x && __double_as_longlong(y) == __double_as_longlong(ry);
      const bool fin = act && (out || cyc || age_now == last);
      const int it = m24 - (int)(last - age_now) * kBlock;  // iterations done at the end of this round
      const bool hit = fin && !out && (cyc || it >= max_it);  // periodic, or ran all max iterations
      const bool back = fin && !hit;                      // escaped in the round, or a short tail
      ws.d_rounds += fin ? (int32_t)age_now : 0;
      ws.d_out += (fin && out) ? 1u : 0u;
      ws.skipped += hit ? (uint32_t)(max_it - it) : 0u;
      ws.n_hit += hit ? 1u : 0u;
      ws.n_cyc += (hit && it < max_it) ? 1u : 0u;
      // an escaped sample goes back to `late` with its round-start state
      if (__ballot_sync(kFull, back))
        push_z<kLate>(q, ws.late_n, back, cx, cy, out ? x0 : x, out ? y0 : y, out ? it - kBlock : it);
      act = act && !fin;
      if constexpr (kCert) {
        // due for the cycle certificate: suspended into the warp's queue in global memory (as many
        // as fit; the others carry on and are due again at the next certificate age)
        const bool due0 = act && cert_on && age_now == ev;
        const unsigned dm = __ballot_sync(kFull, due0);
        if (dm) {
          const int room = (int)p.cert_cap - kCertHeadroom - ws.pend_n - ws.fail_n;
          const int rank = __popc(dm & lanemask_lt());
          const bool due = due0 && rank < room;
          if (due) {
            const size_t idx = (size_t)gwarp * p.cert_cap + (p.cert_cap - 1u - (uint32_t)(ws.pend_n + rank));
            __stcg(p.cert_c + idx, make_double2(cx, cy)); __stcg(p.cert_z + idx, make_double2(x, y));
            __stcg(p.cert_m + idx, make_uint4(last, age_now, 0u, 0u));
          }
          if (due0 && !due) ev = min(last, next_cert_age(age_now));
          ws.pend_n += max(min(__popc(dm), room), 0);
          ws.d_rounds += due ? (int32_t)age_now : 0;
          act = act && !due;
        }
      }
      // with a very large -m a few never-escaping samples could wrap the 32-bit counter
      if (__ballot_sync(kFull, (ws.skipped >> 30) != 0u)) flush_counters(ws, counters);
    }
  }
  ws.d_rounds += act ? (int32_t)age : 0;
  push_deep(q, ws, act, cx, cy, x, y, rx, ry, last, age);  // keeps the checkpoint and its schedule
  __syncwarp();
}

// (c') the cycle certificate over the warp's queue of suspended deep samples (see above).
//
// The work per sample is (Newton passes) x (period) derivative steps, anything from 2 to 320, so a
// batch of 32 in lockstep runs at the pace of its slowest member (measured: 5 passes x ~54 steps
// for every batch, more than the iterations saved on all but config 2).  The samples are therefore
// collected in a queue of p.cert_cap entries per warp in global memory (they are rare: 0.17 % of
// the candidates from age 64, 48 bytes each) and worked off a few hundred at a time:
//   A. period search, 32 entries in lockstep, the period written back into the entry;
//   B. Newton with PER-LANE REFILL: a lane that has settled its sample takes the next entry that
//      has a period, so the cost follows the mean work, not the maximum;
//   C. the entries that got no certificate are compacted to the bottom of the queue, from where
//      reinject_phase returns them to `deep` (32 at a time, whenever deep has room) with their
//      state, age and schedule; the checkpoint restarts at the current state (every certificate
//      age is a checkpoint age anyway).
// Certified samples are `hit max`.
constexpr uint32_t kCertDone = 0xffffffffu;      // meta.z: 0 = no certificate, 1..64 = period found
constexpr double kCertHopeless = 1e-6;           // residual^2 (scaled) a lane must have reached after 2 passes

// position of the n-th (0-based) set bit of m, which has more than n bits set
__device__ __forceinline__ int nth_set_bit(unsigned m, int n) {
  int pos = 0;   // the largest t with at most n set bits below bit t
#pragma unroll
  for (int b = 16; b >= 1; b >>= 1)
    if (__popc(m & ((1u << (pos + b)) - 1u)) <= n) pos += b;
  return pos;
}

// The queues are read and written through L2 (ld.cg / st.cg): an entry is written by one lane and
// read by another lane of the same warp, phases apart.
template <int kVar>
__device__ __forceinline__ void cert_phase(const RenderParams &p, WarpState &ws,
                                           unsigned long long *counters, uint32_t gwarp) {
  const int n = ws.pend_n;                       // entries [cap - n, cap) of the warp's block
  const size_t base = (size_t)gwarp * p.cert_cap + (p.cert_cap - (uint32_t)n);
  const int lane = (int)lane_id();
  __syncwarp();
  // A. period search in float: the first k <= pmax with |z_k - z| < kCertTol
#pragma unroll 1
  for (int c0 = 0; c0 < n; c0 += 32) {
    const bool act = c0 + lane < n;
    double2 c = make_double2(0.0, 0.0), z = make_double2(0.0, 0.0);
    unsigned age = 0;
    if (act) {
      c = __ldcg(p.cert_c + base + c0 + lane); z = __ldcg(p.cert_z + base + c0 + lane);
      age = __ldcg(&p.cert_m[base + c0 + lane].y);
    }
    const int pmax = act ? (age == kCertFirstAge ? kCertPmaxFirst : kCertPmaxLater) : 0;
    const float fcx = (float)c.x, fcy = (float)c.y, fx0 = (float)z.x, fy0 = (float)z.y;
    const float tol2 = 4.0f * kCertTol * kCertTol;   // scaled state: distances are doubled
    float fx = fx0, fy = fy0;
    int per = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < kCertPmaxLater; k0 += 8) {
#pragma unroll
      for (int j = 1; j <= 8; j++) {
        const float a4 = fy * fy, b4 = __fmaf_rn(fx, fx, -a4), yn = __fmaf_rn(fx, fy, fcy);
        fx = __fmaf_rn(b4, 0.5f, fcx); fy = yn;
        const float dx = fx - fx0, dy = fy - fy0;
        if (per == 0 && __fmaf_rn(dy, dy, dx * dx) < tol2) per = k0 + j;
      }
      if (__all_sync(kFull, per != 0 || k0 + 8 >= pmax)) break;
    }
    if (per > pmax) per = 0;
    if (act) __stcg(&p.cert_m[base + c0 + lane].z, (uint32_t)per);
  }
  __syncwarp();
  // B. Newton on f^per(w) - w in double, the derivative d = prod 2 w_j carried along
  {
    const double res2_max = kCertResMax * kCertResMax;
    const int max_it = p.max_it;
    const int m24 = max_it - (max_it - kT2End) % kBlock;
    bool act = false;
    double cx = 0.0, cy = 0.0, zx = 0.0, zy = 0.0, x = 0.0, y = 0.0, dr = 1.0, di = 0.0;
    int per = 0, k = 0, pass = 0, mine = 0, it_done = 0;
    int cursor = 0, chunk0 = 0;
    unsigned avail = 0u;                         // entries of chunk0 .. chunk0+31 with a period, not yet taken
#pragma unroll 1
    for (;;) {
      unsigned idle = __ballot_sync(kFull, !act);
#pragma unroll 1
      while (idle != 0u) {
        if (avail == 0u) {
          if (cursor >= n) break;
          const uint32_t pe = (cursor + lane < n) ? __ldcg(&p.cert_m[base + cursor + lane].z) : 0u;
          avail = __ballot_sync(kFull, pe != 0u);
          chunk0 = cursor; cursor += 32;
          continue;
        }
        const int rank = __popc(idle & lanemask_lt()), na = __popc(avail);
        if (!act && rank < na) {
          mine = chunk0 + nth_set_bit(avail, rank);
          const double2 c = __ldcg(p.cert_c + base + mine), z = __ldcg(p.cert_z + base + mine);
          const uint4 m = __ldcg(p.cert_m + base + mine);               // (last, age, period, -)
          per = (int)m.z;
          it_done = m24 - (int)(m.x - m.y) * kBlock;                    // iterations the sample has done
          cx = c.x; cy = c.y; zx = z.x; zy = z.y; x = zx; y = zy; dr = 1.0; di = 0.0;
          k = 0; pass = 0; act = true;
        }
        // the lowest popc(idle) entries of `avail` are taken now
        avail = __ballot_sync(kFull, ((avail >> lane) & 1u) != 0u &&
                                         __popc(avail & lanemask_lt()) >= __popc(idle));
        idle = __ballot_sync(kFull, !act);
      }
      const unsigned am = __ballot_sync(kFull, act);
      if (am == 0u) break;
      // derivative steps.  A lane that has completed its period waits; the pass-end handling below
      // (~80 issue cycles, ~230 when lanes finish and refill, against 23 per step) runs once
      // kCertGang lanes are waiting, or all of them.
      unsigned waiting;
#pragma unroll 1
      do {
        if (act && k < per) {
          const double ndr = __fma_rn(x, dr, -__dmul_rn(y, di)), ndi = __fma_rn(x, di, __dmul_rn(y, dr));
          dr = ndr; di = ndi;
          zstep<false>(x, y, cx, cy);
          k++;
        }
        waiting = __ballot_sync(kFull, act && k == per);
      } while (__popc(waiting) < kCertGang && waiting != am);
      // end of a pass: accept, give up, or take the Newton step
      const bool e = act && k == per;
      const double rx = __dsub_rn(x, zx), ry = __dsub_rn(y, zy);           // residual (scaled)
      const double res2 = __fma_rn(ry, ry, __dmul_rn(rx, rx));
      const bool conv = res2 < res2_max;
      const bool ok = e && conv && __fma_rn(di, di, __dmul_rn(dr, dr)) < kCertLam2Max;
      const bool giveup = e && !ok && (conv || pass + 1 >= kCertPasses ||
                                       (pass >= 1 && !(res2 < kCertHopeless)));  // (true for NaN)
      if (e && !ok && !giveup) {
        const double er = __dsub_rn(dr, 1.0), ei = di;
        // (an approximate reciprocal is enough: a slightly inexact Newton step still contracts)
        const double inv = (double)__frcp_rn((float)__fma_rn(er, er, __dmul_rn(ei, ei)));
        // w <- w - r / (d - 1) = w - r conj(d - 1) / |d - 1|^2
        zx = __dsub_rn(zx, __dmul_rn(__fma_rn(rx, er, __dmul_rn(ry, ei)), inv));
        zy = __dsub_rn(zy, __dmul_rn(__fma_rn(ry, er, -__dmul_rn(rx, ei)), inv));
        x = zx; y = zy; dr = 1.0; di = 0.0; k = 0; pass++;
      }
      if (__ballot_sync(kFull, ok || giveup)) {
        if (ok) {
          ws.n_hit += 1u; ws.n_cyc += 1u;
          ws.skipped += (uint32_t)(max_it - it_done);
          __stcg(&p.cert_m[base + mine].z, kCertDone);
        } else if (giveup) {
          __stcg(&p.cert_m[base + mine].z, 0u);
        }
        act = act && !(ok || giveup);
        if (__ballot_sync(kFull, (ws.skipped >> 30) != 0u)) flush_counters(ws, counters);  // (huge -m)
      }
    }
  }
  __syncwarp();
  // C. the entries without a certificate move to the bottom of the block.  Chunks are read from
  // the lowest pending index upwards, so the write position (at most the number of entries read
  // before this chunk) stays below everything unread; kCertHeadroom keeps it below the chunk
  // being read as well.
  int f = 0;
#pragma unroll 1
  for (int c0 = 0; c0 < n; c0 += 32) {
    const bool act = c0 + lane < n;
    uint4 m = make_uint4(0u, 0u, kCertDone, 0u);
    double2 c = make_double2(0.0, 0.0), z = make_double2(0.0, 0.0);
    if (act) m = __ldcg(p.cert_m + base + c0 + lane);
    const bool failed = act && m.z == 0u;
    if (failed) { c = __ldcg(p.cert_c + base + c0 + lane); z = __ldcg(p.cert_z + base + c0 + lane); }
    const unsigned fm = __ballot_sync(kFull, failed);   // (also orders the loads before the stores)
    if (failed) {
      const size_t dst = (size_t)gwarp * p.cert_cap + (size_t)(f + __popc(fm & lanemask_lt()));
      __stcg(p.cert_c + dst, c); __stcg(p.cert_z + dst, z); __stcg(p.cert_m + dst, m);
    }
    f += __popc(fm);
  }
  ws.fail_n = f;
  ws.pend_n = 0;
  __syncwarp();
}

// Returns up to 32 samples that got no certificate to `deep` (which holds < 32 entries here).
__device__ __forceinline__ void reinject_phase(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                               uint32_t gwarp) {
  const int take = min(ws.fail_n, 32);
  const bool act = (int)lane_id() < take;
  double2 c = make_double2(0.0, 0.0), z = make_double2(0.0, 0.0);
  uint4 m = make_uint4(0u, 0u, 0u, 0u);
  if (act) {
    const size_t idx = (size_t)gwarp * p.cert_cap + (size_t)(ws.fail_n - 1 - (int)lane_id());
    c = __ldcg(p.cert_c + idx); z = __ldcg(p.cert_z + idx); m = __ldcg(p.cert_m + idx);
  }
  ws.fail_n -= take;
  push_deep(q, ws, act, c.x, c.y, z.x, z.y, z.x, z.y, m.x, m.y);
  __syncwarp();
}

// (d) orbit pass: re-iterate accepted samples for exactly i+1 steps (no escape test needed: the
// count is known), scatter every point with a fire-and-forget red.global.add.u32.
struct OrbitLane {
  bool act;
  double cx, cy, x, y;
  int n;
  uint32_t inc;  // fused render: in-canvas points of this orbit not yet credited to its channels
};

// Division-free binning of one orbit point.  Per axis ONE DFMA gives T = rn(quotient + 2^-11) on a
// 2^-12 grid, on the upper side of the reference's IEEE quotient Q: T - 2^-10 < Q < T (DESIGN.md
// section 4).  If T has left the binade the point is certainly outside; if its fraction is at least
// 2^-10 (any of the bits in kBinClearMask set) then floor(T) = trunc(Q); in the remaining sliver
// (2^-10 of a pixel per axis, exact pixel boundaries included) the point takes bin_exact_index,
// the reference arithmetic verbatim.  (Round 1 formed both T+ and T- = rn(quotient - 2^-11) and
// compared their floors: the same sliver for two more DFMAs and six more integer instructions.)
// The common path is branch-free in the build that never tiles (kDirect: a predicated reduction).
// (fused render: o.n carries the band above bit kOrbStepBits; the point goes to that band's
// histogram)
template <int kVar, bool kDirect>
__device__ __forceinline__ void orbit_bin(const RenderParams &p, OrbitLane &o, WarpState &ws,
                                          const Sink &hist) {
  const double tc = __fma_rn(o.x, p.inv_half_re, p.c0_re);
  const double tr = __fma_rn(o.y, p.inv_half_im, p.c0_im);
  const uint32_t ch = (uint32_t)__double2loint(tc), rh = (uint32_t)__double2loint(tr);
  const bool inrange = ((uint32_t)__double2hiint(tc) == kBinHiWord) &
                       ((uint32_t)__double2hiint(tr) == kBinHiWord);   // (false for the NaN constants)
  const bool clear = ((ch & kBinClearMask) != 0u) & ((rh & kBinClearMask) != 0u);
  const uint32_t col = ch >> kBinFracBits, row = rh >> kBinFracBits;
  // (bitwise on purpose: no short-circuit branches in the common path)
  const bool hit = o.act & inrange & clear & (col < (uint32_t)p.w) & (row < (uint32_t)p.h);
  const bool slow = o.act & ((p.fast_bin == 0) | (inrange & !clear));
  if constexpr ((kVar & kVarFused) != 0) {
    uint32_t idx = row * (uint32_t)p.w + col;
    bool in = hit;
    if (slow) {
      ws.n_exact += 1u;
      in = bin_exact_index(o.x, o.y, p, &idx);
    }
    o.inc += in ? 1u : 0u;
    if (in) scatter<kDirect>(p, hist, idx + ((unsigned)o.n >> kOrbStepBits) * p.band_stride);
  } else {
    if constexpr (kDirect) {
      red_add_u32_if(hist.hist + (row * (uint32_t)p.w + col), hit);
    } else {
      if (hit) scatter<false>(p, hist, row * (uint32_t)p.w + col);
    }
    ws.p_inc += hit ? 1u : 0u;
    if (slow) {
      ws.n_exact += 1u;
      ws.p_inc += bin_exact<kDirect>(o.x, o.y, p, hist) ? 1u : 0u;
    }
  }
}

// Fused render: credit the in-canvas points an orbit has collected to its band (per orbit, not
// per point).  Call before o.n is overwritten or given away.
template <int kVar>
__device__ __forceinline__ void orbit_credit(OrbitLane &o, WarpState &ws) {
  if constexpr ((kVar & kVarFused) != 0) {
    const unsigned band = (unsigned)o.n >> kOrbStepBits;
    ws.p_inc += o.inc;
#pragma unroll
    for (int k = 0; k < kMaxBands; k++) ws.ch_inc[k] += (band == (unsigned)k) ? o.inc : 0u;
    o.inc = 0;
  }
}

template <int kVar, bool kDirect>
__device__ __forceinline__ void orbit_step(const RenderParams &p, OrbitLane &o, WarpState &ws,
                                           const Sink &hist) {
  BUDDHA_ZSTEP(o.x, o.y, o.cx, o.cy);
  orbit_bin<kVar, kDirect>(p, o, ws, hist);
  o.n -= o.act ? 1 : 0;
  o.act = o.act && ((kVar & kVarFused) ? (o.n & ((1 << kOrbStepBits) - 1)) : o.n) != 0;
}

template <int kVar, bool kDirect>
__device__ __forceinline__ void orbit_phase(const RenderParams &p, WarpQueues &q, WarpState &ws,
                                            const Sink &hist) {
  OrbitLane o = {false, 0.0, 0.0, 0.0, 0.0, 0, 0u};
#pragma unroll 1
  for (;;) {
    if (ws.orb_n > 0 && __ballot_sync(kFull, !o.act)) {
      unsigned m = __ballot_sync(kFull, !o.act);
      int rank = __popc(m & lanemask_lt());
      if (!o.act && rank < ws.orb_n) {
        const int slot = slot_of<kOrb, kZCap>(ws.orb_n - 1 - rank);
        orbit_credit<kVar>(o, ws);
        double2 c = q.z_c[slot], z = q.z_z[slot];
        o.cx = c.x; o.cy = c.y; o.x = z.x; o.y = z.y; o.n = q.z_it[slot];
        o.act = true;
      }
      ws.orb_n -= min(__popc(m), ws.orb_n);
      __syncwarp();
    }
    unsigned am = __ballot_sync(kFull, o.act);
    if (__popc(am) < kOrbExit) break;
    orbit_step<kVar, kDirect>(p, o, ws, hist);  // two steps per refill check: orbits are >= 20 steps long on
    orbit_step<kVar, kDirect>(p, o, ws, hist);  // every BASELINE workload, a finished lane idles for one step
  }
  orbit_credit<kVar>(o, ws);
  if (__ballot_sync(kFull, o.act)) push_z<kOrb>(q, ws.orb_n, o.act, o.cx, o.cy, o.x, o.y, o.n);
  __syncwarp();
}

// Orbit entries a warp could not run with enough lanes are spilled to a grid-wide list and
// finished by orbit_drain_kernel, where lanes refill from the whole grid's leftovers.
struct OrbitSpill {
  double4 *entries;          // (cx, cy, x, y)
  int *steps;                // remaining steps
  unsigned int *count;       // [0] entries written, [1] / [2] cursors of the two drain passes,
                             // [3] entries with at least kDrainLong steps left
  unsigned int capacity;
  // Carry-over inside a pipeline of launches (tiled contexts): instead of spilling, a warp parks its
  // < 32 leftover orbits in its own 32 slots and the same warp of the NEXT launch takes them back,
  // so only the last launch of a render call needs the drain.
  double4 *carry_entries;    // [warp * 32 + k]
  int *carry_steps;
  unsigned int *carry_count; // [warp]
  int carry_in, carry_out;
};

// Dynamic shared memory: kWarpsPerCta WarpQueues, then (tiling only) n_tiles list entries per warp.
constexpr size_t kQueueBytes = sizeof(WarpQueues) * kWarpsPerCta;

// Two register budgets (24 warps per SM either way).  kRegsWide = 80: the fastest render when it has
// the SM to itself (+3 % over 72).  kRegsLean = 72 leaves 10240 registers per SM, enough for one
// drain CTA and one apply CTA of the tiled pipeline NEXT TO the resident render CTAs: with 80 the
// side kernels of launch k could only start when launch k+1 had finished (measured: 20000x20000
// -m 20000 3.7e10 -> 4.4e10 samples/s, -m 2000 5.5e10 -> 7.6e10).
constexpr int kRegsWide = 80, kRegsLean = 72;
template <int kVar, int kMaxReg, bool kCertBuild = false>
__global__ void __maxnreg__(kMaxReg)
render_persistent_kernel(RenderParams p, uint32_t *__restrict__ hist,
                         unsigned long long *__restrict__ cursor,
                         unsigned long long *__restrict__ counters, OrbitSpill spill) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpQueues &q = reinterpret_cast<WarpQueues *>(smem_raw)[threadIdx.x >> 5];
  uint2 *tile_tab = reinterpret_cast<uint2 *>(smem_raw + kQueueBytes);
  const Sink sink = {hist + (size_t)(blockIdx.x % p.n_copies) * p.copy_stride,
                     tile_tab + (threadIdx.x >> 5) * p.n_tiles,
                     blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)};
  tile_counters_load(p, sink, true);
  if constexpr ((kVar & kVarFused) != 0) {
    for (int k = (int)lane_id(); k < kChCount; k += 32) q.ch_cnt[k] = 0ull;
    __syncwarp();
  }
  WarpState ws;
  ws.t0_n = ws.t2_n = ws.late_n = ws.deep_n = ws.orb_n = ws.pend_n = ws.fail_n = 0;
  if (spill.carry_in) {  // the orbits this warp parked at the end of the previous launch
    // (the broadcast tells the compiler the count is warp-uniform: without it every vote in the
    //  scheduler loop below is compiled with a divergence fallback, +40 % code, -5 % speed)
    const int n = __shfl_sync(kFull, (int)spill.carry_count[sink.gwarp], 0);
    const int k = (int)lane_id();
    if (k < n) {
      const double4 e = spill.carry_entries[(size_t)sink.gwarp * 32 + k];
      const int slot = slot_of<kOrb, kZCap>(k);
      q.z_c[slot] = make_double2(e.x, e.y); q.z_z[slot] = make_double2(e.z, e.w);
      q.z_it[slot] = spill.carry_steps[(size_t)sink.gwarp * 32 + k];
    }
    ws.orb_n = n;
    __syncwarp();
  }
  ws.chunk_base = 0;
  ws.chunk_off = ws.chunk_len = 0;
  ws.exhausted = false;
  ws.n_rej = ws.n_hit = ws.n_acc = ws.n_cyc = ws.n_exact = 0;
  ws.steps = ws.skipped = ws.d_out = ws.p_pts = ws.p_inc = 0;
  ws.d_rounds = 0;
#pragma unroll
  for (int k = 0; k < kMaxBands; k++) ws.ch_inc[k] = 0;

  // kCertBuild: the build with the cycle certificate (cert_phase).  Like the period-4 test it only
  // pays in the plain 80-register build, and only for large -m, so it is a separate instantiation:
  // contexts that do not use it run the kernel without its code (config 1 and 4 measured 4 % slower
  // with the certificate's code merely present).
#pragma unroll 1
  for (;;) {
    // strict priority along the push graph: a phase is reached only when every stack it pushes
    // to holds < 32 entries.  Once the sample range is used up (`dry`) the partial stacks are run
    // dry, upstream first.
    // `quiet`: no downstream stack needs service and the range is not used up -- the common case
    // (sampler / first tier / tier 2 take turns), decided with two compares instead of the whole
    // chain below (the dispatch was 9 % of the instructions and 14 % of the stall samples of a
    // config-2 launch).  (a | b | c) >= 32 <=> one of them is: all three are < 128.
    const bool quiet = !ws.exhausted && (unsigned)(ws.orb_n | ws.late_n | ws.deep_n) < 32u &&
                       ws.late_n + ws.orb_n <= kZJoint &&
                       (!kCertBuild || (ws.fail_n == 0 && ws.pend_n < (int)p.cert_cap - kCertHeadroom - 64));
    bool dry = false, dry1 = false;
    if (!quiet) {
      dry = ws.exhausted && ws.chunk_off >= ws.chunk_len;
      dry1 = dry && ws.t0_n == 0;
      const bool dry2 = dry1 && ws.t2_n == 0, dry3 = dry2 && ws.late_n == 0;
      if (ws.orb_n >= 32 || (ws.orb_n >= kOrbExit && ws.late_n + ws.orb_n > kZJoint)) {
        orbit_phase<kVar, kMaxReg == kRegsWide>(p, q, ws, sink);  // (tiled contexts take the lean build)
        continue;
      }
      if (ws.late_n >= 32 || ws.late_n + ws.orb_n > kZJoint || (dry2 && ws.late_n > 0)) {
        late_phase<kVar, (kMaxReg == kRegsWide) && (kVar & kVarFused) == 0>(p, q, ws, counters);
        continue;
      }
      if (ws.deep_n >= 32 || (dry3 && ws.deep_n > 0 && ws.fail_n == 0 && ws.pend_n == 0)) {
        // (at the end the parked samples are settled first -- one certificate pass over the queue,
        //  the rest back into deep -- so that everything left drains together)
        deep_phase<kVar, kCertBuild>(p, q, ws, dry3 && ws.fail_n == 0 && ws.pend_n == 0, !dry3,
                                     counters, sink.gwarp);
        continue;
      }
      if (kCertBuild && ws.fail_n > 0) {
        reinject_phase(p, q, ws, sink.gwarp);     // (deep holds < 32 here)
        continue;
      }
      if (kCertBuild && ws.pend_n > 0 &&
          (dry3 || ws.pend_n >= (int)p.cert_cap - kCertHeadroom - 64)) {
        cert_phase<kVar>(p, ws, counters, sink.gwarp);   // (no failed entries left here)
        continue;
      }
    }
    if (ws.t2_n >= 32 || (dry1 && ws.t2_n > 0)) {
      tier_phase<kVar, kT1End, kT2End - kT1End, true>(p, q, ws, ws.t2_n, counters);
    } else if (ws.t0_n >= 32 || (dry && ws.t0_n > 0)) {
      first_tier_phase<kVar>(p, q, ws, counters);
    } else if (!dry) {
      gen_phase<kVar>(p, q, ws, cursor, counters);
    } else {
      break;
    }
  }
  // leftovers (< 32 accepted samples): hand them to the grid-wide list
  if (spill.carry_out) {
    const int k = (int)lane_id();
    if (k < ws.orb_n) {
      const int slot = slot_of<kOrb, kZCap>(k);
      spill.carry_entries[(size_t)sink.gwarp * 32 + k] =
          make_double4(q.z_c[slot].x, q.z_c[slot].y, q.z_z[slot].x, q.z_z[slot].y);
      spill.carry_steps[(size_t)sink.gwarp * 32 + k] = q.z_it[slot];
    }
    if (k == 0) spill.carry_count[sink.gwarp] = (unsigned)ws.orb_n;
  } else if (ws.orb_n > 0) {
    unsigned base = 0;
    if (lane_id() == 0) base = atomicAdd(spill.count, (unsigned)ws.orb_n);
    base = __shfl_sync(kFull, base, 0);
    const int k = (int)lane_id();   // orb_n < 32 here
    bool is_long = false;
    if (k < ws.orb_n && base + k < spill.capacity) {
      const int slot = slot_of<kOrb, kZCap>(k);
      const int n = q.z_it[slot];
      spill.entries[base + k] = make_double4(q.z_c[slot].x, q.z_c[slot].y, q.z_z[slot].x, q.z_z[slot].y);
      spill.steps[base + k] = n;
      is_long = ((kVar & kVarFused) ? (n & ((1 << kOrbStepBits) - 1)) : n) >= kDrainLong;
    }
    const unsigned nl = __popc(__ballot_sync(kFull, is_long));
    if (lane_id() == 0 && nl) atomicAdd(spill.count + 3, nl);
  }
  tile_counters_store(p, sink);
  flush_counters(ws, counters);
  if constexpr ((kVar & kVarFused) != 0) {
    flush_channel_counters<kVar>(p, ws, q.ch_cnt);
    __syncwarp();
    for (int k = (int)lane_id(); k < kMaxBands * kChSlots; k += 32)
      if (q.ch_cnt[k]) atomicAdd(counters + kCntSlots + k, q.ch_cnt[k]);
  }
}

// Finishes the spilled orbits.  The leftovers are few but up to max_it steps long, so their cost is
// latency: a lane stepping its own orbit needs ~360 cycles per recorded point (z-step, binning and
// bookkeeping in one dependent instruction stream) although the z-chain itself is 17.7 cycles per
// step.  So an orbit is given to a GROUP of g lanes (g = 1..32, the largest power of two for which
// all orbits still fit the grid twice over): every lane of the group runs the same z-chain, lane k
// keeps point k of each block of g steps, and the g points are binned at once -- about
// (17.7 g + 300) / g cycles per point.  The redundant FP64 work is irrelevant at this volume.
// Small CTAs (kDrainWarps warps), so that in a pipeline of launches the drain of launch k fits
// next to the resident render CTAs of launch k+1; warp w continues list column w of the render
// kernel (tiling).
constexpr int kDrainWarps = 4;

template <int kVar>
__global__ void __launch_bounds__(kDrainWarps * 32)
orbit_drain_kernel(RenderParams p, uint32_t *__restrict__ hist,
                   unsigned long long *__restrict__ counters, OrbitSpill spill,
                   unsigned int *__restrict__ next, int long_pass) {
  const unsigned total = min(*spill.count, spill.capacity);
  extern __shared__ uint2 tile_tab[];
  const uint32_t gwarp = blockIdx.x * kDrainWarps + (threadIdx.x >> 5);
  if (p.tile_shift && gwarp >= p.n_warps) return;  // no list column to append to
  const Sink sink = {hist, tile_tab + (threadIdx.x >> 5) * p.n_tiles, gwarp};
  tile_counters_load(p, sink, false);  // continue the lists where the render kernel stopped
  WarpState ws;
  ws.n_rej = ws.n_hit = ws.n_acc = ws.n_cyc = ws.n_exact = 0;
  ws.steps = ws.skipped = ws.d_out = ws.p_pts = ws.p_inc = 0;
  ws.d_rounds = 0;
#pragma unroll
  for (int k = 0; k < kMaxBands; k++) ws.ch_inc[k] = 0;

  // Two passes per launch.  A warp steps its 32 / g orbits in lockstep until the longest is done,
  // so a few 20000-step orbits among thousands of short ones keep whole warps spinning with one
  // busy lane group (measured: 2.3e9 warp instructions, 4.3 ms per launch at 20000x20000
  // -m 20000).  The first pass (long_pass) therefore takes only the entries with >= kDrainLong steps
  // left, the second the rest; each with the largest group size g for which its entries still fit
  // the grid twice over.
  const unsigned n_long = min(spill.count[3], total);
  unsigned long long warps = (unsigned long long)gridDim.x * kDrainWarps;
  if (p.tile_shift && warps > p.n_warps) warps = p.n_warps;
  const unsigned long long lanes = 2ull * 32ull * warps;
  const int steps_lo = long_pass ? kDrainLong : 0, steps_hi = long_pass ? 0x7fffffff : kDrainLong;
  const unsigned mine = long_pass ? n_long : total - n_long;
  if (mine == 0u) return;
  int g = 32;
  while (g > 1 && (unsigned long long)mine * g > lanes) g >>= 1;
  const int lane = (int)lane_id();
  const int sub = lane & (g - 1), leader = lane & ~(g - 1);
  double cx = 0.0, cy = 0.0, x = 0.0, y = 0.0;
  int n = 0;          // steps this group's orbit still has to record (group-uniform)
  unsigned mask = 0;  // fused render: its band
  bool more = true;   // the list may still hold entries (group-uniform)
#pragma unroll 1
  for (;;) {
    const unsigned needm = __ballot_sync(kFull, n <= 0 && more && sub == 0);
    if (needm) {
      unsigned base = 0;
      if (lane == 0) base = atomicAdd(next, (unsigned)__popc(needm));
      base = __shfl_sync(kFull, base, 0);
      unsigned idx = base + __popc(needm & lanemask_lt());  // meaningful on the leaders
      idx = __shfl_sync(kFull, idx, leader);
      if (n <= 0 && more) {
        if (idx < total) {
          const double4 e = spill.entries[idx];
          cx = e.x; cy = e.y; x = e.z; y = e.w; n = spill.steps[idx];
          if constexpr ((kVar & kVarFused) != 0) { mask = (unsigned)n >> kOrbStepBits; n &= (1 << kOrbStepBits) - 1; }
          if (n < steps_lo || n >= steps_hi) n = 0;   // the other pass's entry
        } else {
          more = false;
        }
      }
    }
    if (__ballot_sync(kFull, n > 0) == 0u) {
      if (__ballot_sync(kFull, more) == 0u) break;
      continue;  // only entries of the other pass so far: grab again
    }
    OrbitLane o = {sub < n, cx, cy, 0.0, 0.0, (int)(mask << kOrbStepBits) | 1, 0u};
#pragma unroll 1
    for (int k = 0; k < g; k++) {
      BUDDHA_ZSTEP(x, y, cx, cy);
      if (sub == k) { o.x = x; o.y = y; }
    }
    orbit_bin<kVar, false>(p, o, ws, sink);
    orbit_credit<kVar>(o, ws);
    n -= g;
  }
  tile_counters_store(p, sink);
  flush_counters(ws, counters);
  flush_channel_counters<kVar>(p, ws, counters + kCntSlots);
}

// Applies the lists of ONE tile: warp w owns list (t, w).  One launch per tile keeps every
// reduction of the launch inside the same 64 MB of histogram, which stays L2-resident (a single
// launch walking all tiles lets the warps drift apart and measured 66 % L2 misses); list entries
// are read once, coalesced, with a streaming (evict-first) load.  CTAs are small (4 warps, 32
// registers) so that they fit next to the resident render CTAs of the following launch.
constexpr int kApplyWarps = 4;

__global__ void __launch_bounds__(kApplyWarps * 32)
apply_tile_kernel(uint32_t *__restrict__ hist, const uint32_t *__restrict__ tcount,
                  const uint32_t *__restrict__ tile_cap, const uint32_t *__restrict__ tile_base,
                  const uint32_t *__restrict__ pool, int t, uint32_t n_warps, int tile_shift) {
  const uint32_t w = blockIdx.x * kApplyWarps + (threadIdx.x >> 5);
  if (w >= n_warps) return;
  const uint32_t cap = tile_cap[t];
  const uint32_t n = min(tcount[(size_t)t * n_warps + w], cap);
  const uint32_t *src = pool + (tile_base[t] + w * cap);
  uint32_t *tile = hist + ((size_t)t << tile_shift);
  uint32_t i = lane_id();
  for (; i + 96 < n; i += 128) {
    uint32_t a = __ldcs(src + i), b = __ldcs(src + i + 32), c = __ldcs(src + i + 64),
             d = __ldcs(src + i + 96);
    red_add_u32(tile + a); red_add_u32(tile + b); red_add_u32(tile + c); red_add_u32(tile + d);
  }
  for (; i < n; i += 32) red_add_u32(tile + __ldcs(src + i));
}

// Fused contexts: channel = sum of the bands in `bands` (bit b = band b) plus, if given, the
// channel's preloaded counts (buddha_load_histogram); mod 2^32 like the cells themselves.
__global__ void __launch_bounds__(256)
channel_sum_kernel(const uint32_t *__restrict__ hist, const uint32_t *__restrict__ preload,
                   uint32_t *__restrict__ out, size_t cells, unsigned bands) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride) {
    uint32_t v = preload ? preload[i] : 0u;
#pragma unroll
    for (int b = 0; b < kMaxBands; b++)
      if ((bands >> b) & 1u) v += hist[(size_t)b * cells + i];
    out[i] = v;
  }
}

// ---- histogram digest -----------------------------------------------------------------------
//
// A digest that can be formed in parallel, for "N GPUs == 1 GPU" checks on 1.6 GB histograms
// without moving them to the host: cells are taken in blocks of kDigestBlock = 4096; inside a
// block, lane l (0..31) folds cells l, l + 32, l + 64, ... with 64-bit FNV-1a on whole cells
// (h = (h ^ cell) * prime from the FNV offset basis; cells past the end count as 0), the 32 lane
// values are folded the same way in lane order into the block digest, and the host folds the
// block digests in block order.  The tests restate it in numpy (blocked_fnv).
constexpr int kDigestBlock = 4096;
constexpr unsigned long long kFnvBasis = 0xcbf29ce484222325ull, kFnvPrime = 0x100000001b3ull;

__global__ void __launch_bounds__(256)
digest_blocks_kernel(const uint32_t *__restrict__ hist, size_t cells,
                     unsigned long long *__restrict__ block_digest, size_t n_blocks) {
  const size_t warps = ((size_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lane = lane_id();
  for (size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; b < n_blocks; b += warps) {
    const size_t base = b * kDigestBlock;
    unsigned long long h = kFnvBasis;
#pragma unroll 4
    for (int k = 0; k < kDigestBlock / 32; k++) {
      const size_t i = base + (size_t)k * 32 + lane;
      const uint32_t v = i < cells ? __ldg(hist + i) : 0u;
      h = (h ^ v) * kFnvPrime;
    }
    unsigned long long d = kFnvBasis;
    for (int l = 0; l < 32; l++) d = (d ^ __shfl_sync(kFull, h, l)) * kFnvPrime;
    if (lane == 0) block_digest[b] = d;
  }
}

// Privatised copies (RenderParams::n_copies): hist[i] += copies 1 .. n-1, which are cleared.
__global__ void __launch_bounds__(256)
fold_copies_kernel(uint32_t *__restrict__ hist, size_t cells, uint32_t n_copies, size_t stride) {
  size_t step = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += step) {
    uint32_t v = hist[i];
    for (uint32_t k = 1; k < n_copies; k++) {
      v += hist[k * stride + i];
      hist[k * stride + i] = 0u;
    }
    hist[i] = v;
  }
}

// dst[i] += src[i] (mod 2^32): merges counts that were copied to the device into a histogram.
__global__ void __launch_bounds__(256)
add_cells_kernel(uint32_t *__restrict__ dst, const uint32_t *__restrict__ src, size_t cells) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += stride)
    dst[i] += src[i];
}

// ---- tone-map (cudabrot.cu:416-468) ----------------------------------------------------------

// GetLinearColorScale's max search (:430-435).
__global__ void __launch_bounds__(256)
hist_max_kernel(const uint32_t *__restrict__ hist, size_t cells, uint32_t *__restrict__ out_max) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  uint32_t m = 0;
  size_t vec = ((size_t)hist & 15) ? 0 : cells / 4;  // (an odd-sized channel of a fused context)
  const uint4 *h4 = reinterpret_cast<const uint4 *>(hist);
  for (size_t k = i; k < vec; k += stride) {
    uint4 v = __ldg(h4 + k);
    m = max(m, max(max(v.x, v.y), max(v.z, v.w)));
  }
  for (size_t k = vec * 4 + i; k < cells; k += stride) m = max(m, hist[k]);
  m = __reduce_max_sync(kFull, m);
  if (lane_id() == 0 && m) atomicMax(out_max, m);
}

// DoGammaCorrection per pixel (:443-449, :460-466) as a table lookup: lut[c] for c < lut_size is
// the reference's 16-bit value for count c (already byte-swapped when the caller wants big-endian
// output); counts beyond the table are resolved by a search in thr[v] = smallest count whose value
// is >= v (the map is monotone in the count).
__device__ __forceinline__ uint16_t tone_lookup(uint32_t c, const uint16_t *__restrict__ lut,
                                                uint32_t lut_size,
                                                const uint32_t *__restrict__ thr, int swap) {
  if (c < lut_size) return __ldg(lut + c);
  uint32_t lo = 0, hi = 65535;  // largest v with thr[v] <= c
  while (lo < hi) {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (__ldg(thr + mid) <= c) lo = mid; else hi = mid - 1;
  }
  uint16_t v = (uint16_t)lo;
  return swap ? (uint16_t)((v << 8) | (v >> 8)) : v;
}

__global__ void __launch_bounds__(256)
tonemap_kernel(const uint32_t *__restrict__ hist, uint16_t *__restrict__ out, size_t cells,
               const uint16_t *__restrict__ lut, uint32_t lut_size,
               const uint32_t *__restrict__ thr, int swap) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t vec = ((size_t)hist & 15) ? 0 : cells / 4;
  const uint4 *h4 = reinterpret_cast<const uint4 *>(hist);
  uint2 *o4 = reinterpret_cast<uint2 *>(out);
  for (size_t k = i; k < vec; k += stride) {
    uint4 v = __ldg(h4 + k);
    uint32_t a = tone_lookup(v.x, lut, lut_size, thr, swap);
    uint32_t b = tone_lookup(v.y, lut, lut_size, thr, swap);
    uint32_t c = tone_lookup(v.z, lut, lut_size, thr, swap);
    uint32_t d = tone_lookup(v.w, lut, lut_size, thr, swap);
    o4[k] = make_uint2(a | (b << 16), c | (d << 16));
  }
  for (size_t k = vec * 4 + i; k < cells; k += stride)
    out[k] = tone_lookup(hist[k], lut, lut_size, thr, swap);
}

// ---- colour combine -------------------------------------------------------------------------
//
// generate_hires_color_image.sh:61-71 and README.md:176-185 turn three grey renders into one colour
// image with tools outside the reference tree (image_combiner: one render per colour;
// image_combiner_hsl: hue / saturation / lightness).  Here the three tone-mapped planes (native
// byte order, values 0..65535) are combined on the GPU into interleaved 16-bit RGB.
//   mode 0 (RGB): (R, G, B) = (a, b, c)
//   mode 1 (HSL): hue = frac(a / 65535 + hue_adjust), saturation = b / 65535, lightness = c / 65535,
//                 the usual HSL -> RGB, each component rounded to nearest
__global__ void __launch_bounds__(256)
combine_rgb_kernel(const uint16_t *__restrict__ a, const uint16_t *__restrict__ b,
                   const uint16_t *__restrict__ c, uint16_t *__restrict__ rgb, size_t pixels,
                   int mode, double hue_adjust, int swap) {
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels; i += stride) {
    uint32_t r = a[i], g = b[i], bl = c[i];
    if (mode == 1) {
      double h = __dadd_rn(__ddiv_rn((double)r, 65535.0), hue_adjust);
      h = __dsub_rn(h, floor(h));
      const double sat = __ddiv_rn((double)g, 65535.0), lig = __ddiv_rn((double)bl, 65535.0);
      const double chroma = __dmul_rn(__dsub_rn(1.0, fabs(__dsub_rn(__dmul_rn(2.0, lig), 1.0))), sat);
      const double h6 = __dmul_rn(h, 6.0);
      const double x = __dmul_rn(chroma, __dsub_rn(1.0, fabs(__dsub_rn(__dsub_rn(h6, __dmul_rn(2.0, floor(__dmul_rn(h6, 0.5)))), 1.0))));
      const int sector = (int)h6;   // 0..5 (h < 1)
      double r1 = 0.0, g1 = 0.0, b1 = 0.0;
      switch (sector) {
        case 0: r1 = chroma; g1 = x; break;
        case 1: r1 = x; g1 = chroma; break;
        case 2: g1 = chroma; b1 = x; break;
        case 3: g1 = x; b1 = chroma; break;
        case 4: r1 = x; b1 = chroma; break;
        default: r1 = chroma; b1 = x; break;
      }
      const double m = __dsub_rn(lig, __dmul_rn(chroma, 0.5));
      r = (uint32_t)__double2int_rn(fmin(fmax(__dmul_rn(__dadd_rn(r1, m), 65535.0), 0.0), 65535.0));
      g = (uint32_t)__double2int_rn(fmin(fmax(__dmul_rn(__dadd_rn(g1, m), 65535.0), 0.0), 65535.0));
      bl = (uint32_t)__double2int_rn(fmin(fmax(__dmul_rn(__dadd_rn(b1, m), 65535.0), 0.0), 65535.0));
    }
    if (swap) {
      r = ((r << 8) | (r >> 8)) & 0xffffu; g = ((g << 8) | (g >> 8)) & 0xffffu;
      bl = ((bl << 8) | (bl >> 8)) & 0xffffu;
    }
    rgb[3 * i] = (uint16_t)r; rgb[3 * i + 1] = (uint16_t)g; rgb[3 * i + 2] = (uint16_t)bl;
  }
}

// ---- roofline probes --------------------------------------------------------------------------

// Peak FP64-pipe issue rate: 8 independent DFMA chains per thread.
__global__ void __launch_bounds__(256)
probe_dfma_kernel(double *out, int iters, double a, double b) {
  double v0 = threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5,
         v6 = v0 + 6, v7 = v0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      v0 = __fma_rn(v0, a, b); v1 = __fma_rn(v1, a, b); v2 = __fma_rn(v2, a, b);
      v3 = __fma_rn(v3, a, b); v4 = __fma_rn(v4, a, b); v5 = __fma_rn(v5, a, b);
      v6 = __fma_rn(v6, a, b); v7 = __fma_rn(v7, a, b);
    }
  }
  double s = v0 + v1 + v2 + v3 + v4 + v5 + v6 + v7;
  if (s == 12345.678) out[0] = s;  // keeps the chains alive
}

// red.global.add.u32 to uniformly random cells of a `cells`-entry array.
__global__ void __launch_bounds__(256)
probe_red_kernel(uint32_t *buf, unsigned long long cells, int per_thread, uint32_t salt) {
  unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long s = (g + 1) * 0x9E3779B97F4A7C15ull + salt;
#pragma unroll 4
  for (int i = 0; i < per_thread; i++) {
    s ^= s >> 30; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 27; s *= 0x94D049BB133111EBull; s ^= s >> 31;
    unsigned long long idx = __umul64hi(s, cells);  // uniform in [0, cells)
    red_add_u32(buf + idx);
    s += 0x9E3779B97F4A7C15ull;
  }
}

}  // namespace buddha
