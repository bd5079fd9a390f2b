/*
 * buddha_oracle.c -- CPU restatement of the cudabrot hot path.  TEST INFRASTRUCTURE ONLY
 * (see buddha_oracle.h for who may call it and how its parity is pinned).
 *
 * Build: gcc -O2 -ffp-contract=off -mfma -fopenmp -fPIC -shared (oracle/Makefile).
 * -ffp-contract=off is load-bearing: every rounding below is spelled out, and fma() appears
 * exactly where the reference's sm_100a SASS has a DFMA.
 */
#include "buddha_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define FMA(a, b, c) __builtin_fma((a), (b), (c))

/* ---- Philox4x32-10 (curand_philox4x32_x.h:88-91,159-192) -------------------------------- */

static inline void philox_round(uint32_t c[4], uint32_t k0, uint32_t k1) {
  uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
  c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}

void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k0 = key[0], k1 = key[1];
  for (int r = 0; r < 10; r++) {
    philox_round(c, k0, k1);
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(c));
}

/* _curand_uniform_double_hq (curand_uniform.h:101-105) then cudabrot.cu:392-393.
 * Both multiply-adds have exact products, so the device's fused form equals this one. */
static inline double uniform_to_coord(uint32_t x, uint32_t y) {
  uint64_t z = (uint64_t)x ^ ((uint64_t)y << 21);
  double u = FMA((double)z, 0x1p-53, 0x1p-54);
  return FMA(u, 4.0, -2.0);
}

void oracle_sample(uint64_t seed, uint64_t s, double *c_real, double *c_imag) {
  uint32_t ctr[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0u, 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t o[4];
  oracle_philox4x32_10(ctr, key, o);
  *c_real = uniform_to_coord(o[0], o[1]);
  *c_imag = uniform_to_coord(o[2], o[3]);
}

/* ---- canvas (cudabrot.cu:505-527) -------------------------------------------------------- */

int oracle_set_deltas(oracle_dims *d) {
  if (d->w <= 0) return 0;
  if (d->h <= 0) return 0;
  if (d->max_real <= d->min_real) return 0;
  if (d->max_imag <= d->min_imag) return 0;
  d->delta_imag = (d->max_imag - d->min_imag) / ((double)d->h);
  d->delta_real = (d->max_real - d->min_real) / ((double)d->w);
  return 1;
}

/* ---- rejection (cudabrot.cu:284-298) ----------------------------------------------------- */

int oracle_rejected(double real, double imag) {
  double i2 = imag * imag;              /* DMUL, shared by both tests */
  double q0 = real - 0.25;              /* DADD */
  double q = FMA(q0, q0, i2);           /* DFMA  :288 */
  double lhs = q * (q0 + q);            /* DADD, DMUL :289 (q + (real-0.25) is commutative) */
  double rhs = i2 * 0.25;               /* DMUL */
  if (lhs < rhs) return 1;
  double t = real + 1.0;                /* DADD :295 */
  double b = FMA(t, t, i2);             /* DFMA :296-297 */
  return b < 0.0625;
}

/* ---- the recurrence (cudabrot.cu:331-336 / :357-363) ------------------------------------- */

#define STEP(re, im, cre, cim)                                  \
  do {                                                          \
    double t1_ = (im) * (im);              /* DMUL          */  \
    double t2_ = FMA((re), (re), -t1_);    /* DFMA          */  \
    double r2_ = (re) + (re);              /* DADD          */  \
    double nre_ = (cre) + t2_;             /* DADD          */  \
    (im) = FMA(r2_, (im), (cim));          /* DFMA          */  \
    (re) = nre_;                                                \
  } while (0)

#define ESCAPED(re, im) (FMA((im), (im), (re) * (re)) > 4.0) /* DMUL, DFMA, DSETP */

/* RENDER_BURNING_SHIP (cudabrot.cu:15-17, :327-330, :353-356): real = fabs(real), imag =
 * fabs(imag) before the step.  SASS of that build (sm_100a, CUDA 12.9): DMUL im,im; DADD |re|,|re|;
 * DFMA re,re,-t1; DFMA r2,|im|,c_im; DADD c_re,t2 -- the same contraction, |.| as operand
 * modifiers. */
#define STEP_SHIP(re, im, cre, cim)                             \
  do {                                                          \
    (re) = fabs(re); (im) = fabs(im);                           \
    STEP(re, im, cre, cim);                                     \
  } while (0)

int oracle_escape_iterations(double c_real, double c_imag, int max_iterations) {
  double re = c_real, im = c_imag;
  for (int i = 0; i < max_iterations; i++) {
    STEP(re, im, c_real, c_imag);
    if (ESCAPED(re, im)) return i;
  }
  return max_iterations;
}

int oracle_escape_iterations_ship(double c_real, double c_imag, int max_iterations) {
  double re = c_real, im = c_imag;
  for (int i = 0; i < max_iterations; i++) {
    STEP_SHIP(re, im, c_real, c_imag);
    if (ESCAPED(re, im)) return i;
  }
  return max_iterations;
}

int oracle_cycle_detect_iterations(double c_real, double c_imag, int max_iterations, int stride) {
  double re = c_real, im = c_imag;
  double ref_re = re, ref_im = im;
  if (stride < 1) stride = 1;
  int next_ckpt = stride;
  for (int i = 0; i < max_iterations; i++) {
    STEP(re, im, c_real, c_imag);
    if (ESCAPED(re, im)) return -1;
    int n = i + 1; /* iterations done */
    if (n % stride == 0) {
      if (memcmp(&re, &ref_re, 8) == 0 && memcmp(&im, &ref_im, 8) == 0) return n;
      if (n == next_ckpt) { ref_re = re; ref_im = im; next_ckpt *= 2; }
    }
  }
  return -1;
}

/* cvt.rzi.s32.f64: truncate toward zero, saturate, NaN -> 0 (host (int) is UB out of range). */
static inline int32_t sat_trunc_i32(double v) {
  if (v != v) return 0;
  if (v >= 2147483648.0) return INT32_MAX;
  if (v <= -2147483649.0) return INT32_MIN;
  return (int32_t)v;
}

/* IncrementPixelCounter, cudabrot.cu:302-314.  Returns 1 if a cell was incremented. */
static inline int bin_point(double re, double im, const oracle_dims *d, int64_t *index) {
  if ((re < d->min_real) || (im < d->min_imag)) return 0;
  int32_t col = sat_trunc_i32((re - d->min_real) / d->delta_real);
  int32_t row = sat_trunc_i32((im - d->min_imag) / d->delta_imag);
  if ((row >= 0) && (row < d->h) && (col >= 0) && (col < d->w)) {
    *index = (int64_t)((int32_t)(row * d->w) + col); /* 32-bit int index, :312 */
    return 1;
  }
  return 0;
}

typedef struct {
  uint32_t *hist;
  int atomic;
} sink_t;

static inline void sink_add(sink_t *s, int64_t idx) {
  if (s->atomic) {
#pragma omp atomic
    s->hist[idx] += 1u;
  } else {
    s->hist[idx] += 1u;
  }
}

static void render_range(const oracle_dims *d, int max_it, int min_it, uint64_t seed,
                         uint64_t first, uint64_t count, sink_t *sink, oracle_counters *c,
                         int ship) {
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    c->candidates++;
    /* cudabrot.cu:397-399: no cardioid/bulb test in the burning-ship build */
    if (!ship && oracle_rejected(cre, cim)) { c->rejected++; continue; }
    int i = ship ? oracle_escape_iterations_ship(cre, cim, max_it)
                 : oracle_escape_iterations(cre, cim, max_it);
    if (i >= max_it) { c->hit_max++; c->escape_iters += (uint64_t)(max_it > 0 ? max_it : 0); continue; }
    c->escape_iters += (uint64_t)i + 1u;
    if (i < min_it) { c->too_early++; continue; }
    c->accepted++;
    /* IterateAndRecord, cudabrot.cu:347-365 */
    double re = cre, im = cim;
    for (;;) {
      if (ship) STEP_SHIP(re, im, cre, cim); else STEP(re, im, cre, cim);
      c->orbit_points++;
      int64_t idx;
      if (bin_point(re, im, d, &idx)) { sink_add(sink, idx); c->increments++; }
      if (ESCAPED(re, im)) break;
    }
  }
}

static void counters_add(oracle_counters *a, const oracle_counters *b) {
  a->candidates += b->candidates; a->rejected += b->rejected; a->hit_max += b->hit_max;
  a->too_early += b->too_early; a->accepted += b->accepted; a->escape_iters += b->escape_iters;
  a->orbit_points += b->orbit_points; a->increments += b->increments;
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

int oracle_render(const oracle_dims *d, int max_iterations, int min_iterations, uint64_t seed,
                  uint64_t first, uint64_t count, uint32_t *hist, oracle_counters *counters,
                  int threads) {
  return oracle_render_ex(d, max_iterations, min_iterations, seed, first, count, hist, counters,
                          threads, 0);
}

int oracle_render_ex(const oracle_dims *d, int max_iterations, int min_iterations, uint64_t seed,
                     uint64_t first, uint64_t count, uint32_t *hist, oracle_counters *counters,
                     int threads, int burning_ship) {
  int nt = threads > 0 ? threads : oracle_max_threads();
  size_t cells = (size_t)d->w * (size_t)d->h;
  oracle_counters total;
  memset(&total, 0, sizeof(total));
  /* private histograms while they stay under ~2 GiB in total, shared + omp atomic beyond */
  int use_private = (nt > 1) && (cells * 4u * (size_t)nt <= ((size_t)2 << 30));
  const uint64_t chunk = 4096;
  uint64_t nchunks = (count + chunk - 1) / chunk;
#ifdef _OPENMP
#pragma omp parallel num_threads(nt)
#endif
  {
    oracle_counters local;
    memset(&local, 0, sizeof(local));
    sink_t sink;
    sink.hist = hist;
    sink.atomic = (nt > 1) && !use_private;
    uint32_t *priv = NULL;
    if (use_private) {
      priv = (uint32_t *)calloc(cells, sizeof(uint32_t));
      if (priv) { sink.hist = priv; sink.atomic = 0; } else { sink.atomic = 1; }
    }
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 1)
#endif
    for (uint64_t ci = 0; ci < nchunks; ci++) {
      uint64_t lo = ci * chunk;
      uint64_t n = (count - lo < chunk) ? (count - lo) : chunk;
      render_range(d, max_iterations, min_iterations, seed, first + lo, n, &sink, &local,
                   burning_ship);
    }
#ifdef _OPENMP
#pragma omp critical
#endif
    {
      counters_add(&total, &local);
      if (priv) {
        for (size_t p = 0; p < cells; p++) hist[p] += priv[p];
      }
    }
    free(priv);
  }
  if (counters) *counters = total;
  return nt;
}

void oracle_classify_ship(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                          int32_t *out_iters) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    out_iters[k] = oracle_escape_iterations_ship(cre, cim, max_iterations);
  }
}

void oracle_classify(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                     int32_t *out_iters) {
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024)
#endif
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    out_iters[k] = oracle_rejected(cre, cim) ? -1
                                             : oracle_escape_iterations(cre, cim, max_iterations);
  }
}

/* ---- tone-map (cudabrot.cu:416-468), host arithmetic: no FMA, glibc pow ------------------ */

static inline uint16_t clamp_u16(double v) { /* Clamp, :416-420 */
  if (v <= 0) return 0;
  if (v >= 0xffff) return 0xffff;
  return (uint16_t)v;
}

/* (uint16_t) of a double as x86-64 gcc does it: cvttsd2si to 32 bits, keep the low 16.
 * NaN / out of range -> 0x80000000 -> 0.  Spelled out so the oracle has no UB. */
static inline uint16_t x86_double_to_u16(double v) {
  if (!(v > -2147483649.0 && v < 2147483648.0)) return 0;
  return (uint16_t)(uint32_t)(int32_t)v;
}

void oracle_tonemap(const uint32_t *hist, size_t cells, double gamma, int big_endian,
                    uint16_t *out, uint32_t *max_out, double *scale_out) {
  uint32_t max = 0;
  for (size_t p = 0; p < cells; p++) if (hist[p] > max) max = hist[p]; /* :430-435 */
  double scale = ((double)0xffff) / ((double)max);                     /* :436 */
  if (max_out) *max_out = max;
  if (scale_out) *scale_out = scale;
  const double m = 0xffff;
  for (size_t p = 0; p < cells; p++) {
    double scaled = ((double)hist[p]) * scale;                         /* :445 */
    uint16_t v;
    if (gamma <= 0.0) {
      v = x86_double_to_u16(scaled);                                   /* :447 */
    } else {
      double g = m * pow(scaled / m, 1 / gamma);                       /* :448 */
      v = (g != g) ? x86_double_to_u16(g) : clamp_u16(g);
    }
    if (big_endian) v = (uint16_t)(((v & 0xff) << 8) | (v >> 8));      /* :568 */
    out[p] = v;
  }
}

int oracle_write_pgm(const char *path, const uint16_t *image, int w, int h) {
  FILE *f = fopen(path, "wb");
  if (!f) return 1;
  if (fprintf(f, "P5\n%d %d\n%d\n", w, h, 0xffff) <= 0) { fclose(f); return 2; }
  size_t n = (size_t)w * (size_t)h;
  uint16_t *tmp = (uint16_t *)malloc(n * 2);
  if (!tmp) { fclose(f); return 3; }
  for (size_t i = 0; i < n; i++) tmp[i] = (uint16_t)(((image[i] & 0xff) << 8) | (image[i] >> 8));
  int ok = fwrite(tmp, n * 2, 1, f) == 1;
  free(tmp);
  fclose(f);
  return ok ? 0 : 4;
}

uint64_t oracle_fnv1a64(const uint32_t *data, size_t cells) {
  uint64_t hsh = 0xcbf29ce484222325ull;
  const uint8_t *b = (const uint8_t *)data;
  for (size_t i = 0; i < cells * 4; i++) { hsh ^= b[i]; hsh *= 0x100000001b3ull; }
  return hsh;
}

/* ---- restatements of the product's transformed arithmetic (see header) ------------------- */

int oracle_escape_iterations_scaled(double c_real, double c_imag, int max_iterations) {
  double cx = c_real * 2.0, cy = c_imag * 2.0, x = cx, y = cy;
  for (int i = 0; i < max_iterations; i++) {
    double a4 = y * y;
    double b4 = FMA(x, x, -a4);
    double yn = FMA(x, y, cy);
    x = FMA(b4, 0.5, cx);
    y = yn;
    if (FMA(y, y, x * x) > 16.0) return i;
  }
  return max_iterations;
}

int oracle_escape_iterations_scaled_ship(double c_real, double c_imag, int max_iterations) {
  double cx = c_real * 2.0, cy = c_imag * 2.0, x = cx, y = cy;
  for (int i = 0; i < max_iterations; i++) {
    double a4 = y * y;
    double b4 = FMA(x, x, -a4);
    double yn = FMA(fabs(x), fabs(y), cy);
    x = FMA(b4, 0.5, cx);
    y = yn;
    if (FMA(y, y, x * x) > 16.0) return i;
  }
  return max_iterations;
}

int oracle_rejected_scaled(double c_real, double c_imag) {
  double cx = c_real * 2.0, cy = c_imag * 2.0;
  double i2 = cy * cy;
  double q0 = cx + -0.5;
  double q = FMA(q0, q0, i2);
  double s = FMA(q0, 2.0, q);
  double lhs = q * s;
  double t = cx + 2.0;
  double b = FMA(t, t, i2);
  return (lhs < i2) || (b < 0.25);
}

int oracle_bin_reference(const oracle_dims *d, double re, double im, int64_t *index) {
  return bin_point(re, im, d, index);
}

typedef struct { double inv_half, c0; int ok; } fast_axis;

/* make_fast_bin of csrc/buddha_api.cu */
static fast_axis make_fast_axis(double min_v, double delta, int n) {
  fast_axis f = {0, 0, 0};
  double inv = 1.0 / delta;
  if (!(inv > 0.0) || !isfinite(inv)) return f;
  if (n > (1 << 19)) return f;
  if (!(fabs(min_v) * inv < 0x1p36)) return f;
  long double base = (long double)0x1.8p40 - (long double)min_v * (long double)inv;
  f.c0 = (double)(base + (long double)0x1p-11);
  f.inv_half = inv * 0.5;
  f.ok = 1;
  return f;
}

static inline uint32_t hi32(double v) { uint64_t b; memcpy(&b, &v, 8); return (uint32_t)(b >> 32); }
static inline uint32_t lo32(double v) { uint64_t b; memcpy(&b, &v, 8); return (uint32_t)b; }

int oracle_bin_fast(const oracle_dims *d, double re, double im, int64_t *index, int *took_exact) {
  fast_axis fr = make_fast_axis(d->min_real, d->delta_real, d->w);
  fast_axis fi = make_fast_axis(d->min_imag, d->delta_imag, d->h);
  *took_exact = 0;
  if (!fr.ok || !fi.ok) return -1;
  double x2 = re * 2.0, y2 = im * 2.0;
  /* orbit_bin of csrc/buddha_kernels.cuh: one rounding per axis on the upper side of the
   * quotient, T - 2^-10 < Q < T; outside the binade = outside the canvas; a fraction below 2^-10
   * (none of the bits 0xffc set) = too close to a pixel boundary: the reference arithmetic decides */
  double tc = FMA(x2, fr.inv_half, fr.c0), tr = FMA(y2, fi.inv_half, fi.c0);
  const uint32_t H0 = 0x42780000u;
  if (hi32(tc) != H0 || hi32(tr) != H0) return 0;
  uint32_t ch = lo32(tc), rh = lo32(tr);
  int clear = ((ch & 0xffcu) != 0u) && ((rh & 0xffcu) != 0u);
  if (!clear) {
    *took_exact = 1;
    return bin_point(x2 * 0.5, y2 * 0.5, d, index);
  }
  uint32_t col = ch >> 12, row = rh >> 12;
  if (col < (uint32_t)d->w && row < (uint32_t)d->h) {
    *index = (int64_t)(row * (uint32_t)d->w + col);
    return 1;
  }
  return 0;
}

uint64_t oracle_check_scaled(uint64_t seed, uint64_t first, uint64_t count, int max_iterations) {
  uint64_t bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : bad)
#endif
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    int r1 = oracle_rejected(cre, cim), r2 = oracle_rejected_scaled(cre, cim);
    if (r1 != r2) { bad++; continue; }
    if (r1) continue;
    if (oracle_escape_iterations(cre, cim, max_iterations) !=
        oracle_escape_iterations_scaled(cre, cim, max_iterations)) bad++;
  }
  return bad;
}

uint64_t oracle_check_scaled_ship(uint64_t seed, uint64_t first, uint64_t count,
                                  int max_iterations) {
  uint64_t bad = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : bad)
#endif
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    if (oracle_escape_iterations_ship(cre, cim, max_iterations) !=
        oracle_escape_iterations_scaled_ship(cre, cim, max_iterations)) bad++;
  }
  return bad;
}

uint64_t oracle_check_fast_bin(const oracle_dims *d, const double *points, uint64_t n,
                               uint64_t *exact_count, uint64_t *in_canvas) {
  uint64_t bad = 0, ex = 0, in = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : bad, ex, in)
#endif
  for (uint64_t k = 0; k < n; k++) {
    int64_t i1 = -1, i2 = -1;
    int took = 0;
    int a = oracle_bin_reference(d, points[2 * k], points[2 * k + 1], &i1);
    int b = oracle_bin_fast(d, points[2 * k], points[2 * k + 1], &i2, &took);
    if (b < 0) { bad++; continue; }
    if (a != b || (a && i1 != i2)) bad++;
    ex += (uint64_t)took;
    in += (uint64_t)a;
  }
  if (exact_count) *exact_count = ex;
  if (in_canvas) *in_canvas = in;
  return bad;
}

/* The kernel's conservative period-3 test (buddha_kernels.cuh: in_period3_component), restated in
 * float arithmetic: lambda = 8 + 4c -+ 4c sqrt(-7 - 4c) from the Giarrusso-Fisher relation
 * c^3 + 2c^2 + (1 - lambda/8)c + (1 - lambda/8)^2 = 0; 1 if min |lambda|^2 < limit. */
int oracle_period3_flag(double c_real, double c_imag, float limit) {
  const float a = (float)c_real, b = (float)c_imag;
  const float wr = fmaf(-4.0f, a, -7.0f), wi = -4.0f * b;
  const float mw = sqrtf(fmaf(wr, wr, wi * wi));
  const float sr = sqrtf(fmaxf(0.5f * (mw + wr), 0.0f));
  const float si = copysignf(sqrtf(fmaxf(0.5f * (mw - wr), 0.0f)), wi);
  const float tr = fmaf(a, sr, -b * si), ti = fmaf(a, si, b * sr);
  const float br = fmaf(4.0f, a, 8.0f), bi = 4.0f * b;
  const float l1r = fmaf(-4.0f, tr, br), l1i = fmaf(-4.0f, ti, bi);
  const float l2r = fmaf(4.0f, tr, br), l2i = fmaf(4.0f, ti, bi);
  const float m1 = fmaf(l1r, l1r, l1i * l1i), m2 = fmaf(l2r, l2r, l2i * l2i);
  return fminf(m1, m2) < limit;
}

/* Evidence for the claim the kernel relies on: every sample the period-3 test flags runs the
 * reference's escape loop (cudabrot.cu:319-340) to max_iterations.  Returns the number of flagged
 * samples that escaped (must be 0); *flagged / *inset count the flagged samples and all samples
 * that hit max_iterations after passing the cardioid / bulb test. */
uint64_t oracle_check_period3(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                              float limit, uint64_t *flagged, uint64_t *inset) {
  uint64_t bad = 0, fl = 0, in = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : bad, fl, in)
#endif
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    if (oracle_rejected(cre, cim)) continue;
    int f = oracle_period3_flag(cre, cim, limit);
    if (!f && inset == NULL) continue;
    int it = oracle_escape_iterations(cre, cim, max_iterations);
    if (it >= max_iterations) in++;
    if (f) { fl++; if (it < max_iterations) bad++; }
  }
  if (flagged) *flagged = fl;
  if (inset) *inset = in;
  return bad;
}

/* The sampler's FP32 pre-classification (buddha_kernels.cuh: prefilter), restated: coordinates from
 * the top 23 bits of the two high Philox words, cardioid / bulb and the first two steps in float,
 * each decision only when it clears its threshold by a margin.  0 = undecided (goes on to the exact
 * FP64 path), 1 = certainly rejected, 2 = certainly escapes at step 1, 3 = certainly at step 2. */
int oracle_prefilter_class(uint32_t hi_re, uint32_t hi_im, int ship, float m_rej, float m_esc) {
  union { uint32_t u; float f; } a, b;
  a.u = 0x3F800000u | (hi_re >> 9);
  b.u = 0x3F800000u | (hi_im >> 9);
  const float cx = fmaf(a.f, 8.0f, -12.0f), cy = fmaf(b.f, 8.0f, -12.0f);   /* 2 * c */
  const float i2 = cy * cy;
  if (!ship) {
    const float q0 = cx - 0.5f, q = fmaf(q0, q0, i2), s = fmaf(q0, 2.0f, q), lhs = q * s;
    const float t = cx + 2.0f, bb = fmaf(t, t, i2);
    if (lhs < i2 - m_rej || bb < 0.25f - m_rej) return 1;
  }
  const float x1 = fmaf(fmaf(cx, cx, -i2), 0.5f, cx);
  const float y1 = ship ? fmaf(fabsf(cx), fabsf(cy), cy) : fmaf(cx, cy, cy);
  const float n1 = fmaf(y1, y1, x1 * x1);
  if (n1 > 16.0f + m_esc) return 2;
  if (!(n1 < 16.0f - m_esc)) return 0;
  const float a2 = y1 * y1;
  const float x2 = fmaf(fmaf(x1, x1, -a2), 0.5f, cx);
  const float y2 = ship ? fmaf(fabsf(x1), fabsf(y1), cy) : fmaf(x1, y1, cy);
  const float n2 = fmaf(y2, y2, x2 * x2);
  if (n2 > 16.0f + m_esc) return 3;
  return 0;
}

/* Evidence for the pre-classification: every decided sample must agree with the reference's own
 * arithmetic (rejected <=> oracle_rejected; "escapes at step k" <=> not rejected and
 * IterateMandelbrot returns k-1).  Returns the number of disagreements (must be 0); counts[4] =
 * samples per class. */
uint64_t oracle_check_prefilter(uint64_t seed, uint64_t first, uint64_t count, int ship,
                                float m_rej, float m_esc, uint64_t counts[4]) {
  uint64_t bad = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static) reduction(+ : bad, c0, c1, c2, c3)
#endif
  for (uint64_t k = 0; k < count; k++) {
    const uint64_t s = first + k;
    uint32_t ctr[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0u, 0u};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    uint32_t o[4];
    oracle_philox4x32_10(ctr, key, o);
    const int cls = oracle_prefilter_class(o[1], o[3], ship, m_rej, m_esc);
    if (cls == 0) { c0++; continue; }
    const double cre = uniform_to_coord(o[0], o[1]), cim = uniform_to_coord(o[2], o[3]);
    const int rej = ship ? 0 : oracle_rejected(cre, cim);
    if (cls == 1) { c1++; if (!rej) bad++; continue; }
    const int it = ship ? oracle_escape_iterations_ship(cre, cim, 3)
                        : oracle_escape_iterations(cre, cim, 3);
    if (cls == 2) { c2++; if (rej || it != 0) bad++; }
    else { c3++; if (rej || it != 1) bad++; }
  }
  if (counts) { counts[0] = c0; counts[1] = c1; counts[2] = c2; counts[3] = c3; }
  return bad;
}

/* The kernel's period-4 test (buddha_kernels.cuh: in_period4_component; used by the 80-register
 * build) restated in float, two Newton steps like the kernel: Newton
 * from 0 on mu^3 - (3 - c^2) mu^2 + (3 + c^2 - c^3 - c^4) mu - (1 + 2c^2 + 3c^3 + 3c^4 + 3c^5 + c^6)
 * (mu = lambda / 16), accepted if |mu| + 3 |p / p'| < mu_max. */
typedef struct { float r, i; } cfl;
static inline cfl cfl_mul(cfl a, cfl b) { cfl o = {fmaf(a.r, b.r, -a.i * b.i), fmaf(a.r, b.i, a.i * b.r)}; return o; }
static inline cfl cfl_add(cfl a, cfl b) { cfl o = {a.r + b.r, a.i + b.i}; return o; }
static inline cfl cfl_div(cfl a, cfl b) {
  const float d = 1.0f / fmaf(b.r, b.r, b.i * b.i);
  cfl o = {fmaf(a.r, b.r, a.i * b.i) * d, fmaf(a.i, b.r, -a.r * b.i) * d};
  return o;
}
int oracle_period4_flag(double c_real, double c_imag, float mu_max) {
  const cfl c = {(float)c_real, (float)c_imag};
  const cfl c2 = cfl_mul(c, c), c3 = cfl_mul(c2, c), c4 = cfl_mul(c2, c2), c5 = cfl_mul(c4, c), c6 = cfl_mul(c3, c3);
  const cfl a2 = {c2.r - 3.0f, c2.i};
  const cfl a1 = {3.0f + c2.r - c3.r - c4.r, c2.i - c3.i - c4.i};
  const cfl a0 = {-(1.0f + 2.0f * c2.r + 3.0f * (c3.r + c4.r + c5.r) + c6.r),
                  -(2.0f * c2.i + 3.0f * (c3.i + c4.i + c5.i) + c6.i)};
  cfl mu = {0.0f, 0.0f}, p = a0, dp = a1;
  for (int k = 0; k < 2; k++) {
    const cfl q = cfl_div(p, dp);
    mu.r -= q.r; mu.i -= q.i;
    p = cfl_add(cfl_mul(cfl_add(cfl_mul(cfl_add(mu, a2), mu), a1), mu), a0);
    cfl t3 = {3.0f * mu.r, 3.0f * mu.i}, t2 = {2.0f * a2.r, 2.0f * a2.i};
    dp = cfl_add(cfl_mul(cfl_add(t3, t2), mu), a1);
  }
  const cfl q = cfl_div(p, dp);
  const float bound = sqrtf(fmaf(mu.r, mu.r, mu.i * mu.i)) + 3.0f * sqrtf(fmaf(q.r, q.r, q.i * q.i));
  return bound < mu_max;
}

/* Like oracle_check_period3, for the period-4 test: flagged samples that escaped (must be 0). */
uint64_t oracle_check_period4(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                              float mu_max, uint64_t *flagged, uint64_t *inset) {
  uint64_t bad = 0, fl = 0, in = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : bad, fl, in)
#endif
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    if (oracle_rejected(cre, cim)) continue;
    int f = oracle_period4_flag(cre, cim, mu_max);
    if (!f && inset == NULL) continue;
    int it = oracle_escape_iterations(cre, cim, max_iterations);
    if (it >= max_iterations) in++;
    if (f) { fl++; if (it < max_iterations) bad++; }
  }
  if (flagged) *flagged = fl;
  if (inset) *inset = in;
  return bad;
}

/* ---- attracting-cycle certificate (buddha_kernels.cuh: cert_phase) ----------------------------
 * Restated on the scaled state (X = 2 re, Y = 2 im) like the kernel:
 *   1. period search in float: the first k <= pmax with |z_k - z_0| < tol (scaled: 2 tol);
 *   2. `passes` Newton steps on f^p(z) - z in double, the derivative d = prod 2 z_j carried along
 *      (on the scaled state 2 z = X + iY, so d <- (X + iY) d); the reciprocal of |d - 1|^2 is
 *      taken in float (a slightly inexact Newton step still contracts);
 *   3. at the last iterate: residual |f^p(z) - z| (scaled) below res_max and |d|^2 < lam2_max.
 * f_c then has an attracting cycle; a quadratic polynomial has at most one, and it attracts the
 * critical orbit (Fatou), so c lies in a hyperbolic component and the orbit of 0 never escapes.
 * Returns the period, or 0 if no certificate was obtained. */
int oracle_cycle_certificate(double c_real, double c_imag, double z_real, double z_imag, int pmax,
                             float tol, int passes, double res_max, double lam2_max) {
  const double cx = 2.0 * c_real, cy = 2.0 * c_imag;
  double zx = 2.0 * z_real, zy = 2.0 * z_imag;
  const float fcx = (float)cx, fcy = (float)cy, fx0 = (float)zx, fy0 = (float)zy;
  float fx = fx0, fy = fy0;
  const float tol2 = 4.0f * tol * tol;
  int p = 0;
  for (int k = 1; k <= pmax; k++) {
    const float a4 = fy * fy, b4 = fmaf(fx, fx, -a4), yn = fmaf(fx, fy, fcy);
    fx = fmaf(b4, 0.5f, fcx); fy = yn;
    const float dx = fx - fx0, dy = fy - fy0;
    if (fmaf(dy, dy, dx * dx) < tol2) { p = k; break; }
  }
  if (!p) return 0;
  double rx = 0.0, ry = 0.0, dr = 1.0, di = 0.0;
  for (int pass = 0; pass < passes; pass++) {
    double x = zx, y = zy;
    dr = 1.0; di = 0.0;
    for (int k = 0; k < p; k++) {
      const double ndr = FMA(x, dr, -(y * di)), ndi = FMA(x, di, y * dr);
      dr = ndr; di = ndi;
      const double a4 = y * y, b4 = FMA(x, x, -a4), yn = FMA(x, y, cy);
      x = FMA(b4, 0.5, cx); y = yn;
    }
    rx = x - zx; ry = y - zy;                      /* scaled residual */
    if (pass + 1 == passes) break;
    const double er = dr - 1.0, ei = di;
    const double inv = (double)(1.0f / (float)FMA(er, er, ei * ei));
    /* z <- z - r / (d - 1) = z - r conj(d - 1) / |d - 1|^2 */
    zx -= FMA(rx, er, ry * ei) * inv;
    zy -= FMA(ry, er, -(rx * ei)) * inv;
  }
  if (!(FMA(ry, ry, rx * rx) < res_max * res_max)) return 0;
  if (!(FMA(di, di, dr * dr) < lam2_max)) return 0;
  return p;
}

/* Evidence for the certificate, in the setting the kernel uses it: samples that survive
 * 46 steps enter rounds of 24 unchecked steps; at ages (rounds) trig0, trig0 * trig_mul, ... the
 * certificate is tried (pmax1 at the first age, pmax2 later; the kernel: 4, 16, 64, ...).  Returns the number of certified samples that ESCAPE
 * under the reference's loop within max_iterations (must be 0).  stats[0] = candidates not
 * rejected, [1] = never-escaping samples not flagged by the period-3/4 tests, [2] = certified among
 * them, [3] = their iterations with bit-exact detection only, [4] = with the certificate,
 * [5] = certificate attempts, [6] = sum over attempts of (pmax searched), [7] = sum over
 * attempts that found a period of that period. */
uint64_t oracle_check_certificate(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                                  int trig0, int trig_mul, int pmax1, int pmax2, float tol, int passes, double res_max,
                                  double lam2_max, uint64_t stats[8]) {
  uint64_t bad = 0, s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0, s7 = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : bad, s0, s1, s2, s3, s4, s5, s6, s7)
#endif
  for (uint64_t k = 0; k < count; k++) {
    double cre, cim;
    oracle_sample(seed, first + k, &cre, &cim);
    if (oracle_rejected(cre, cim)) continue;
    s0++;
    const int it = oracle_escape_iterations(cre, cim, max_iterations);
    if (it < 46) continue;
    if (oracle_period3_flag(cre, cim, 0.998f) || oracle_period4_flag(cre, cim, 0.999f / 16.0f)) continue;
    double re = cre, im = cim;
    for (int j = 0; j < 46; j++) {
      double t1 = im * im, t2 = FMA(re, re, -t1), r2 = re + re;
      im = FMA(r2, im, cim); re = cre + t2;
    }
    double rre = re, rim = im;
    unsigned age = 0, trig = (unsigned)trig0;
    int done = 46, t_exact = -1, t_cert = -1;
    while (done + 24 <= max_iterations) {
      if (age == trig) {
        trig *= (unsigned)trig_mul;
        if (t_cert < 0) {
          const int pm = age == (unsigned)trig0 ? pmax1 : pmax2;
          const int p = oracle_cycle_certificate(cre, cim, re, im, pm, tol, passes, res_max, lam2_max);
          s5++; s6 += (uint64_t)pm; s7 += (uint64_t)p;
          if (p) t_cert = done;
        }
      }
      { const unsigned rest = age & (age - 1u); if (rest == 0u || 3u * rest == 2u * age) { rre = re; rim = im; } }
      for (int j = 0; j < 24; j++) {
        double t1 = im * im, t2 = FMA(re, re, -t1), r2 = re + re;
        im = FMA(r2, im, cim); re = cre + t2;
      }
      done += 24; age++;
      if (!(FMA(im, im, re * re) <= 4.0)) break;
      if (memcmp(&re, &rre, 8) == 0 && memcmp(&im, &rim, 8) == 0) { t_exact = done; break; }
    }
    if (it < max_iterations) { if (t_cert >= 0) bad++; continue; }
    s1++;
    const int te = t_exact >= 0 ? t_exact : max_iterations;
    s3 += (uint64_t)te;
    if (t_cert >= 0) { s2++; s4 += (uint64_t)t_cert; } else s4 += (uint64_t)te;
  }
  if (stats) { stats[0] = s0; stats[1] = s1; stats[2] = s2; stats[3] = s3; stats[4] = s4; stats[5] = s5; stats[6] = s6; stats[7] = s7; }
  return bad;
}
