/*
 * buddha_oracle.h -- CPU restatement of the cudabrot hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is the checker, never the product: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path (cudabrot_b200/csrc) never
 * links or calls anything in oracle/.
 *
 * Parity status: the reference (/root/reference/cudabrot.cu) ships no tests, golden vectors or
 * CPU path, so parity is pinned by
 *   (1) Random123 Philox4x32-10 known-answer vectors (tests/test_oracle.py),
 *   (2) the reference's own device functions executed on a B200 over this oracle's sample list
 *       (oracle/ref_probe.cu includes the reference source where it lies; `-m gpu` test), and
 *   (3) the reference's own host tone-map/PGM code executed here on the CPU (same ref_probe binary),
 *       with the outputs committed under tests/golden/.
 */
#ifndef BUDDHA_ORACLE_H
#define BUDDHA_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors FractalDimensions, cudabrot.cu:46-58 (same field order, 56 bytes). */
typedef struct {
  int32_t w, h;
  double min_real, min_imag, max_real, max_imag;
  double delta_real, delta_imag;
} oracle_dims;

/* Work counters: S, E, P, I and the classification of every candidate. */
typedef struct {
  uint64_t candidates;   /* S: every drawn c, including rejected ones                         */
  uint64_t rejected;     /* in main cardioid or period-2 bulb (cudabrot.cu:398)               */
  uint64_t hit_max;      /* ran max iterations without escaping (cudabrot.cu:407)             */
  uint64_t too_early;    /* escaped with i < min (cudabrot.cu:408)                            */
  uint64_t accepted;     /* min <= i < max                                                    */
  uint64_t escape_iters; /* E: iterations IterateMandelbrot executes (i+1, or max)            */
  uint64_t orbit_points; /* P: sum of (i+1) over accepted samples                             */
  uint64_t increments;   /* I: orbit points that landed inside the canvas                     */
} oracle_counters;

/* Philox4x32-10, /usr/local/cuda/include/curand_philox4x32_x.h:88-91,159-192. */
void oracle_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);

/* Sample index s -> candidate c (replaces cudabrot.cu:392-393 with the stateless Philox stream). */
void oracle_sample(uint64_t seed, uint64_t s, double *c_real, double *c_imag);

/* RecomputePixelDeltas, cudabrot.cu:505-527.  Returns 1 if valid (deltas filled in), else 0. */
int oracle_set_deltas(oracle_dims *d);

/* InMainCardioid || InOrder2Bulb, cudabrot.cu:284-298 (SASS dataflow). */
int oracle_rejected(double real, double imag);

/* IterateMandelbrot, cudabrot.cu:319-340. */
int oracle_escape_iterations(double c_real, double c_imag, int max_iterations);

/* First bit-exact repeat of z (orbit provably periodic in this arithmetic): returns the iteration
 * index at which Brent's algorithm sees z == checkpoint, or -1 if none within max_iterations.
 * Used only to study / test the product's exact non-escape shortcut. */
int oracle_cycle_detect_iterations(double c_real, double c_imag, int max_iterations, int stride);

/* DrawBuddhabrot over samples [first, first+count), race-free (every increment counts).
 * hist is uint32[h*w], accumulated into (not cleared).  threads<=0 -> omp_get_max_threads().
 * Returns the number of OpenMP threads actually used. */
int oracle_render(const oracle_dims *d, int max_iterations, int min_iterations, uint64_t seed,
                  uint64_t first, uint64_t count, uint32_t *hist, oracle_counters *counters,
                  int threads);

/* Same with burning_ship != 0: the RENDER_BURNING_SHIP build of the reference (cudabrot.cu:15-17,
 * :327-330, :353-356, :397-399): |re|, |im| before every step, no cardioid/bulb test. */
int oracle_render_ex(const oracle_dims *d, int max_iterations, int min_iterations, uint64_t seed,
                     uint64_t first, uint64_t count, uint32_t *hist, oracle_counters *counters,
                     int threads, int burning_ship);
int oracle_escape_iterations_ship(double c_real, double c_imag, int max_iterations);
void oracle_classify_ship(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                          int32_t *out_iters);
int oracle_escape_iterations_scaled_ship(double c_real, double c_imag, int max_iterations);
uint64_t oracle_check_scaled_ship(uint64_t seed, uint64_t first, uint64_t count,
                                  int max_iterations);

/* Escape classification only (no histogram): out_iters[k] = IterateMandelbrot result for sample
 * first+k, or -1 if rejected by the cardioid/bulb test. */
void oracle_classify(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                     int32_t *out_iters);

/* SetGrayscalePixels + GetLinearColorScale + DoGammaCorrection + Clamp, cudabrot.cu:416-468.
 * big_endian!=0 additionally applies SaveImage's byte swap (cudabrot.cu:566-570). */
void oracle_tonemap(const uint32_t *hist, size_t cells, double gamma, int big_endian,
                    uint16_t *out, uint32_t *max_out, double *scale_out);

/* SaveImage, cudabrot.cu:548-577: "P5\n<w> <h>\n65535\n" + big-endian pixels.  image is
 * host-endian on entry (not modified).  Returns 0 on success. */
int oracle_write_pgm(const char *path, const uint16_t *image, int w, int h);

/* FNV-1a-64 over the little-endian bytes of a uint32 array (the survey's histogram digest). */
uint64_t oracle_fnv1a64(const uint32_t *data, size_t cells);

int oracle_max_threads(void);

/* ---- CPU restatements of the PRODUCT's transformed arithmetic --------------------------------
 * The kernels do not run the reference dataflow literally: they keep the orbit scaled by two
 * (4 FP64 instructions per step) and bin without divisions.  These functions restate those two
 * transformations on the CPU so tests can compare them with the reference-form functions above
 * over millions of ordinary and adversarial inputs without a GPU. */

/* Escape index computed with the scaled recurrence of csrc/buddha_kernels.cuh (BUDDHA_ZSTEP). */
int oracle_escape_iterations_scaled(double c_real, double c_imag, int max_iterations);

/* cardioid/bulb test on the scaled candidate (rejected2 in buddha_kernels.cuh). */
int oracle_rejected_scaled(double c_real, double c_imag);

/* Reference binning (cudabrot.cu:302-314): returns 1 and the cell index if (re, im) is counted. */
int oracle_bin_reference(const oracle_dims *d, double re, double im, int64_t *index);

/* Division-free binning of orbit_bin (one DFMA rounding per axis, on the upper side of the quotient); *took_exact is set when the
 * fast path declined and the IEEE-division path decided.  Returns -1 if the canvas does not
 * admit the fast path at all (product falls back to exact binning for every point). */
int oracle_bin_fast(const oracle_dims *d, double re, double im, int64_t *index, int *took_exact);

/* Batch checkers for the two functions above (OpenMP).  Both return the number of mismatches
 * against the reference-form functions (0 = identical everywhere). */
uint64_t oracle_check_scaled(uint64_t seed, uint64_t first, uint64_t count, int max_iterations);
/* points = n (re, im) pairs; *exact_count receives how many took the division path. */
uint64_t oracle_check_fast_bin(const oracle_dims *d, const double *points, uint64_t n,
                               uint64_t *exact_count, uint64_t *in_canvas);

/* The kernel's conservative period-3 membership test restated in float arithmetic, and the check
 * that every sample it flags runs the reference's escape loop to max_iterations (returns the number
 * that did not: must be 0). */
int oracle_period3_flag(double c_real, double c_imag, float limit);
uint64_t oracle_check_period3(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                              float limit, uint64_t *flagged, uint64_t *inset);

/* The sampler's FP32 pre-classification restated (0 undecided, 1 rejected, 2 / 3 escapes at step
 * 1 / 2), and the check that every decided sample agrees with the reference's arithmetic (returns
 * the disagreements: must be 0; counts[4] = samples per class). */
int oracle_prefilter_class(uint32_t hi_re, uint32_t hi_im, int ship, float m_rej, float m_esc);
uint64_t oracle_check_prefilter(uint64_t seed, uint64_t first, uint64_t count, int ship,
                                float m_rej, float m_esc, uint64_t counts[4]);

/* The same pair for the period-4 test (Newton for a small root of the multiplier cubic). */
int oracle_period4_flag(double c_real, double c_imag, float mu_max);
uint64_t oracle_check_period4(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                              float mu_max, uint64_t *flagged, uint64_t *inset);

/* The kernel's attracting-cycle certificate (buddha_kernels.cuh: cert_phase) restated: float period
 * search from the state z, Newton on f^p(z) - z in double, accepted if the residual and the
 * multiplier are small.  Returns the period or 0.  oracle_check_certificate tries it where the
 * kernel does (deep ages 4, 16, 64, ...) and returns how many certified samples escape under the
 * reference's loop (must be 0); stats[8] is described at the definition. */
int oracle_cycle_certificate(double c_real, double c_imag, double z_real, double z_imag, int pmax,
                             float tol, int passes, double res_max, double lam2_max);
uint64_t oracle_check_certificate(uint64_t seed, uint64_t first, uint64_t count, int max_iterations,
                                  int trig0, int trig_mul, int pmax1, int pmax2, float tol, int passes, double res_max,
                                  double lam2_max, uint64_t stats[8]);

#ifdef __cplusplus
}
#endif
#endif
