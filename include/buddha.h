/*
 * buddha.h -- C ABI of libbuddha.so, the B200-native replacement for cudabrot's hot path.
 *
 * The reference (/root/reference/cudabrot.cu) has no plugin or FFI surface: host and device code
 * share one translation unit and meet at two <<<>>> launches and four cudaMem* calls.  Each export
 * below names the reference call site it replaces (file:line into /root/reference/).  Everything
 * is plain C: pointers, sizes and PODs; no CUDA, torch or C++ types cross the boundary.
 *
 * Conventions: every call returns 0 on success or a BUDDHA_E* code; nothing here ever calls
 * exit() (the reference's CheckCUDAError, cudabrot.cu:30,134-141, prints and exits -- the CLI in
 * csrc/cudabrot_main.c keeps that behaviour on top of these return codes).  A context is bound to
 * one GPU and must be used from one host thread at a time.  There is no CPU fallback: if no
 * sm_100 device is usable, buddha_create fails.
 */
#ifndef BUDDHA_H
#define BUDDHA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BUDDHA_ABI_VERSION 3
#define BUDDHA_MAX_CHANNELS 4

enum {
  BUDDHA_OK = 0,
  BUDDHA_EINVAL = 1,   /* bad argument / invalid canvas (cudabrot.cu:505-527 rules)   */
  BUDDHA_ECUDA = 2,    /* a CUDA runtime call failed; see buddha_last_error           */
  BUDDHA_ENOMEM = 3,
  BUDDHA_ESIZE = 4,    /* buffer size does not match w*h (cudabrot.cu:239-245)        */
  BUDDHA_ENCCL = 5,    /* NCCL not loadable / collective failed                       */
  BUDDHA_ENODEV = 6    /* no usable sm_100 GPU                                        */
};

/* buddha_params.flags */
#define BUDDHA_F_NO_SHORTCUT   (1u << 0) /* disable the exact periodicity shortcut (same output)   */
#define BUDDHA_F_SIMPLE_KERNEL (1u << 1) /* one-sample-per-thread debug kernel, reference dataflow */
#define BUDDHA_F_EXACT_BINNING (1u << 2) /* always bin with IEEE divisions (same output)           */
#define BUDDHA_F_FORCE_TILED   (1u << 3) /* tile-binned scatter even for small histograms (tests)  */
#define BUDDHA_F_BURNING_SHIP  (1u << 4) /* RENDER_BURNING_SHIP (cudabrot.cu:15-17, :327-330,
                                            :353-356, :397-399): |re|, |im| before every step, no
                                            cardioid/bulb test; a compile-time switch there     */

/* Canvas + iteration limits: FractalDimensions (cudabrot.cu:46-58) and IterationControl (:62-67),
 * plus what the reference hard-codes (seed, :37) or keeps in its global struct (device, :72). */
typedef struct {
  uint32_t struct_size;     /* = sizeof(buddha_params), for ABI evolution */
  int32_t device;           /* -d, cudabrot.cu:667-671 */
  int32_t width, height;    /* -w / -h */
  double min_real, max_real, min_imag, max_imag; /* --min-real ... --max-imag */
  int32_t max_iterations;   /* -m */
  int32_t min_iterations;   /* -c */
  uint64_t seed;            /* DEFAULT_RNG_SEED = 1337 in the reference */
  uint32_t flags;
  uint32_t reserved;
  /* Fused multi-channel render (no reference equivalent: generate_hires_color_image.sh:27-59 runs
   * the reference once per colour channel).  n_channels >= 2 renders every candidate ONCE and adds
   * its orbit to each channel k whose window channel_min[k] <= i < channel_max[k] accepts it;
   * max_iterations / min_iterations above are then ignored.  The histogram becomes
   * uint32[n_channels][h][w] at this API (on the device it is one histogram per distinct set of
   * accepting channels, which is what buddha_device_histogram points to); channel k is bit-identical to what a single-channel context with
   * (-m channel_max[k], -c channel_min[k]) renders from the same sample indices.  Every
   * channel_max must be > 22 and < 2^28.  0 or 1 = the plain single-channel render. */
  uint32_t n_channels;
  int32_t channel_max[BUDDHA_MAX_CHANNELS];
  int32_t channel_min[BUDDHA_MAX_CHANNELS];
  uint32_t reserved2;
} buddha_params;

/* Work counters, accumulated over every render call since create / reset_counters.
 * S = candidates, E = escape_iters, P = orbit_points, I = increments. */
typedef struct {
  uint64_t candidates;     /* every drawn c, rejected ones included                          */
  uint64_t rejected;       /* main cardioid / period-2 bulb (cudabrot.cu:398)                */
  uint64_t hit_max;        /* never escaped within max_iterations (:407)                     */
  uint64_t too_early;      /* escaped before min_iterations (:408)                           */
  uint64_t accepted;       /* orbit recorded                                                 */
  uint64_t escape_iters;   /* iterations the reference's IterateMandelbrot would execute     */
  uint64_t orbit_points;   /* recorded steps = sum(i+1) over accepted                        */
  uint64_t increments;     /* orbit points that landed inside the canvas                     */
  uint64_t executed_iters; /* escape-pass iterations this library actually executed          */
  uint64_t shortcut_hits;  /* hit_max samples proven periodic before reaching max_iterations */
  uint64_t kernel_launches;/* render + tone-map kernels launched by this context             */
  uint64_t exact_bins;     /* orbit points that took the IEEE-division binning path          */
} buddha_counters;

typedef struct buddha_ctx buddha_ctx;

uint32_t buddha_abi_version(void);

/* Default parameters = the reference's defaults (cudabrot.cu:763-772, :530-543):
 * 1000x1000, [-2,2]^2, -m 100, -c 20, device 0, seed 1337. */
void buddha_default_params(buddha_params *p);

/* RecomputePixelDeltas (cudabrot.cu:505-527): validates and returns the pixel spacing exactly as
 * the reference computes it, (max-min)/(double)dim.  0 if valid, BUDDHA_EINVAL otherwise; *why (if
 * not NULL) receives the reference's own message for the first failed rule. */
int buddha_validate_canvas(const buddha_params *p, double *delta_real, double *delta_imag,
                           const char **why);

/* SetupCUDA (cudabrot.cu:153-189): select the device, allocate and zero the uint32 histogram.
 * There is no RNG state to allocate: sample s of seed k is Philox4x32-10(ctr=(s,0,0), key=k). */
int buddha_create(buddha_ctx **out, const buddha_params *p);

/* CleanupGlobals (cudabrot.cu:112-119). */
void buddha_destroy(buddha_ctx *ctx);

/* Last error text for this context ("" if none).  buddha_create failures: pass NULL. */
const char *buddha_last_error(const buddha_ctx *ctx);

/* cudaMemset of the histogram (cudabrot.cu:169). */
int buddha_clear_histogram(buddha_ctx *ctx);

/* H2D copy of a -s file's content (cudabrot.cu:256-257).  cells must equal w*h. */
int buddha_load_histogram(buddha_ctx *ctx, const uint32_t *host, size_t cells);

/* D2H copy of the histogram (cudabrot.cu:496-497); row-major uint32[h][w], row 0 = min_imag.
 * Fused contexts: cells = n_channels*w*h for load/read (channel-major); buddha_read_channel copies
 * one channel (cells = w*h), i.e. exactly one -s file of the reference. */
int buddha_read_histogram(buddha_ctx *ctx, uint32_t *host, size_t cells);
int buddha_read_channel(buddha_ctx *ctx, int channel, uint32_t *host, size_t cells);

/* The pass loop + DrawBuddhabrot (cudabrot.cu:483-492, :379-414), reproducible form: renders
 * exactly the candidates with sample indices [first, first+count) into the histogram and returns
 * when the GPU is done.  Any split of an index range over calls, contexts or GPUs gives the same
 * summed histogram. */
int buddha_render_samples(buddha_ctx *ctx, uint64_t first, uint64_t count);

/* Same, but only enqueues the work on the context's stream; pair with buddha_sync. */
int buddha_render_samples_async(buddha_ctx *ctx, uint64_t first, uint64_t count);
int buddha_sync(buddha_ctx *ctx);

/* -t semantics (cudabrot.cu:483-492): render consecutive sample indices starting at `first` until
 * `seconds` have elapsed (checked between passes, so seconds == 0 still renders one pass; negative
 * = until *stop becomes nonzero) or *stop (may be NULL) is set, e.g. by a SIGINT handler (:756-760).
 * *samples_done receives the number of candidates rendered, *passes the number of launches. */
int buddha_render_seconds(buddha_ctx *ctx, double seconds, volatile int *stop, uint64_t first,
                          uint64_t *samples_done, uint64_t *passes);

/* Device time (CUDA events on the context's stream) of the most recent render call, in ms. */
int buddha_last_render_ms(buddha_ctx *ctx, float *ms);

int buddha_get_counters(buddha_ctx *ctx, buddha_counters *out);
/* Fused contexts: the counters a single-channel render with channel k's limits would report
 * (candidates, rejected, hit_max, too_early, accepted, escape_iters, orbit_points, increments);
 * executed_iters, shortcut_hits, exact_bins and kernel_launches are those of the fused pass.
 * buddha_get_counters itself then describes the fused pass: limits = the widest channel,
 * accepted = samples accepted by at least one channel, increments = cells hit, once per point. */
int buddha_get_channel_counters(buddha_ctx *ctx, int channel, buddha_counters *out);
int buddha_reset_counters(buddha_ctx *ctx);

/* SetGrayscalePixels (cudabrot.cu:454-468 with :425-439, :443-449, :416-420) and, when big_endian
 * is nonzero, SaveImage's byte swap (:566-570), on the GPU.  host_out receives w*h uint16 (may be
 * NULL to only get max/scale).  *max_out / *scale_out are the two numbers the reference prints
 * ("Max value: %lu, scale: %f", :437).  Bit-identical to the reference's host code, glibc pow
 * included (the count->grey map is tabulated on the host with the reference's expression). */
int buddha_tonemap_u16(buddha_ctx *ctx, double gamma, int big_endian, uint16_t *host_out,
                       size_t cells, uint32_t *max_out, double *scale_out);

/* Same for one channel of a fused context (channel 0 of a plain one): each channel is scaled by
 * its own maximum, as three separate runs of the reference would. */
int buddha_tonemap_channel_u16(buddha_ctx *ctx, int channel, double gamma, int big_endian,
                               uint16_t *host_out, size_t cells, uint32_t *max_out,
                               double *scale_out);

/* Device time of the most recent tone-map kernels (max-reduce + map), in ms. */
int buddha_last_tonemap_ms(buddha_ctx *ctx, float *ms);

/* Overlapped host transfers.  The reference's copies (H2D cudabrot.cu:256-257, D2H :496-497) block
 * and its tone-map runs on the host while the GPU idles (:500); these calls let a long job stream
 * results out, and saved counts in, while the next sample range renders:
 *   buddha_add_histogram_async  ADDS host counts (a saved -s buffer, cells as for load_histogram)
 *       to the histogram: the copy runs on its own stream beside whatever is rendering, the add is
 *       ordered after the work enqueued so far.  `host` must stay valid until buddha_sync (use
 *       page-locked memory for a truly asynchronous copy).
 *   buddha_snapshot             freezes the histogram as of all work enqueued so far into a second
 *       device buffer (returns at once).
 *   buddha_read_snapshot / buddha_tonemap_snapshot_u16   the same results as buddha_read_histogram /
 *       buddha_tonemap_channel_u16 would have given at the snapshot, delivered on a copy stream:
 *       they block the caller, but neither wait for nor delay renders enqueued after the snapshot. */
int buddha_add_histogram_async(buddha_ctx *ctx, const uint32_t *host, size_t cells);
int buddha_snapshot(buddha_ctx *ctx);
int buddha_read_snapshot(buddha_ctx *ctx, uint32_t *host, size_t cells);
int buddha_tonemap_snapshot_u16(buddha_ctx *ctx, int channel, double gamma, int big_endian,
                                uint16_t *host_out, size_t cells, uint32_t *max_out,
                                double *scale_out);

/* Colour image from three channels, on the GPU.  generate_hires_color_image.sh:61-71 and
 * README.md:176-185 do this with tools outside the reference tree (ImageMagick, image_combiner,
 * image_combiner_hsl), so there is no byte-level reference; the definition is:
 * each channels[k] (an index into this context's channels; indices may repeat) is tone-mapped with
 * `gamma` and scaled by its own maximum exactly as buddha_tonemap_channel_u16 does, then
 *   BUDDHA_COMBINE_RGB: (R, G, B) = the three grey values
 *   BUDDHA_COMBINE_HSL: hue = frac(g0 / 65535 + hue_adjust), saturation = g1 / 65535, lightness =
 *                       g2 / 65535, standard HSL -> RGB in double precision, rounded to nearest.
 * host_rgb receives pixels * 3 uint16 (interleaved; big_endian != 0: PPM "P6 ... 65535" byte order).
 * max_out (may be NULL) receives the three maxima. */
#define BUDDHA_COMBINE_RGB 0
#define BUDDHA_COMBINE_HSL 1
int buddha_combine_rgb_u16(buddha_ctx *ctx, const int channels[3], double gamma, int mode,
                           double hue_adjust, int big_endian, uint16_t *host_rgb, size_t pixels,
                           uint32_t max_out[3]);

/* 64-bit digest of one channel of the histogram (channel 0 of a plain context), formed on the GPU:
 * equal histograms give equal digests, so "N GPUs == 1 GPU == CPU restatement" can be checked on 1.6 GB
 * canvases without moving them.  Definition (restated in numpy by the tests, blocked_fnv):
 * blocks of 4096 cells; in a block, lane l = 0..31 folds cells l, l+32, ... with FNV-1a-64 taken
 * over whole cells (h = (h ^ cell) * 0x100000001b3 from 0xcbf29ce484222325; cells past the end are
 * 0); the 32 lane values are folded the same way into the block digest, the block digests into the
 * result. */
int buddha_histogram_digest(buddha_ctx *ctx, int channel, uint64_t *digest);

/* Multi-GPU plumbing (no reference equivalent; cudabrot is single-GPU).
 * Raw device pointer + CUDA stream of this context, so a host framework (torch.distributed / NCCL)
 * can reduce the private histograms in place. */
void *buddha_device_histogram(buddha_ctx *ctx);
/* Cells behind that pointer: w*h, or for a fused context w*h times its number of bands. */
size_t buddha_device_histogram_cells(buddha_ctx *ctx);
void *buddha_stream(buddha_ctx *ctx);

/* In-process merge: sums the histograms of n contexts (one per GPU) into ctxs[root] with a single
 * ncclReduce(ncclUint32, ncclSum) per GPU; NCCL is dlopen()ed on first use. */
int buddha_merge(buddha_ctx **ctxs, int n, int root);

/* Roofline probes, measured on this context's GPU:
 * peak FP64-pipe issue rate in lane-instructions/s (independent DFMA chains), and
 * red.global.add.u32 throughput in ops/s to uniformly random cells of a footprint_bytes array. */
int buddha_probe_fp64_peak(buddha_ctx *ctx, double *lane_instr_per_s);
int buddha_probe_red_peak(buddha_ctx *ctx, size_t footprint_bytes, double *red_per_s);

#ifdef __cplusplus
}
#endif
#endif /* BUDDHA_H */
