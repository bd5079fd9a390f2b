// buddha_api.cu -- host side of libbuddha.so: the C ABI declared in include/buddha.h.
//
// Each export cites the reference call site it replaces in include/buddha.h.  This file owns the
// device memory, the stream and the launch policy; the kernels are in buddha_kernels.cuh.
// There is no CPU fallback anywhere in this library.
//
// Environment switches (tuning / experiments; none changes a result):
//   BUDDHA_TILE_MIN_MB        histogram size from which the tile-binned scatter is used (640)
//   BUDDHA_TILE_SHIFT         log2(cells per tile) (24 = 64 MB)
//   BUDDHA_TILE_POOL_MB       size of the list pool (min(16 GB, free/4))
//   BUDDHA_TILE_LAUNCH_LOG2   log2 of the largest launch of the tiled pipeline (30)
//   BUDDHA_TILE_SERIAL        apply / drain on the render stream (needed under ncu kernel replay)
//   BUDDHA_PAD_SMEM           extra dynamic shared memory per CTA (occupancy experiments)
//   BUDDHA_NO_CARRY           drain the leftover orbits after every pipeline launch (no carry-over)
//   BUDDHA_TILE_TRACE         print device timestamps of every pipeline launch to stderr
//   BUDDHA_CERT_QUEUE         entries per warp of the cycle-certificate queues (512; 0 = no certificate)
#include "../../include/buddha.h"
#include "buddha_kernels.cuh"

#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <vector>

using namespace buddha;

namespace {

thread_local char g_create_error[512] = "";

constexpr uint32_t kLutMax = 1u << 18;  // counts below this are tone-mapped through a full table
constexpr int kTotalCnt = kCntSlots + kMaxBands * kChSlots;  // common + per-channel / per-band accumulators

struct FastBin {
  double inv_half, c0;
  bool ok;
};

}  // namespace

struct buddha_ctx {
  buddha_params params;
  double delta_re, delta_im;
  size_t cells;                   // all channels
  size_t ch_cells;                // w * h
  size_t dev_cells;               // cells of the device histogram: ch_cells x (bands, or 1)
  int n_ch;                       // 1, or the channels of a fused context
  int n_bands;                    // fused: distinct channel sets along the step axis (RenderParams)
  unsigned band_set[kMaxBands];   // fused: channels (bit k) of each band
  uint32_t *d_preload;            // fused: counts given to buddha_load_histogram, [n_ch][h][w]
  int sm_count;
  int grid;                       // persistent grid: resident CTAs per SM x SMs
  int variant;                    // kernel instantiation: kVarShip | kVarFused
  cudaStream_t stream;
  cudaEvent_t ev_a, ev_b, ev_ta, ev_tb;
  uint32_t *d_hist;
  unsigned long long *d_cursor;   // offset handed out so far in the current launch
  unsigned long long *d_counters; // kCntSlots accumulators
  OrbitSpill spill[2];            // grid-wide lists of orbits left over by the render kernel
  unsigned int *d_spill_next[2];  // (two: the drain of launch k overlaps the render of launch k+1)
  double2 *d_cert_c, *d_cert_z;   // certificate queues (cert_phase in buddha_kernels.cuh), kCertQueue
  uint4 *d_cert_m;                // entries per warp of the render grid; only where the build uses them
  // tile-binned scatter (histograms far beyond L2), see scatter() in buddha_kernels.cuh
  bool tiled, tile_calibrated;
  int tile_shift, n_tiles;
  uint32_t tile_warps;            // warps of the full grid = lists per tile
  size_t n_lists, tile_smem;
  size_t render_smem, smem_pad;      // dynamic shared memory of the render kernel (stacks [+ tile counters])
  uint32_t *d_tcount, *d_tcap, *d_pool;   // d_tcount / d_pool hold two buffers each; d_tcap per tile
  cudaStream_t apply_stream;              // apply_tiles_kernel of launch k overlaps render k+1
  cudaEvent_t ev_rendered, ev_applied[2];
  bool apply_pending[2];
  int tile_buf;
  uint32_t *d_tbase;              // per tile: first pool entry (inside one pool half)
  size_t pool_entries;
  double tile_pts_per_sample;
  // tone-map scratch: [0] for the live histogram (render stream), [1] for the snapshot (copy stream)
  struct ToneBufs {
    uint32_t *d_max;
    uint16_t *d_gray;             // tone-mapped image, allocated on first use
    uint32_t *d_chan;             // fused contexts: one assembled channel, allocated on first use
    uint16_t *d_lut;
    uint32_t *d_thr;
    uint32_t lut_capacity;
  } tone[2];
  // asynchronous host transfers (buddha_add_histogram_async, buddha_snapshot, ...)
  cudaStream_t h2d_stream, d2h_stream;
  cudaEvent_t ev_staged, ev_added, ev_snap;
  uint32_t *d_stage;              // counts on their way into the histogram
  uint32_t *d_snap;               // snapshot of the device histogram (all bands)
  uint32_t *d_snap_preload;       // fused: snapshot of the loaded counts
  bool snap_valid;
  uint16_t *d_planes;             // colour combine: three grey planes + the RGB image
  unsigned long long *d_digest;   // block digests (buddha_histogram_digest)
  size_t digest_capacity;
  RenderParams rp;
  uint64_t candidates;            // host tally: every index in a rendered range is a candidate
  uint64_t launches;
  bool render_timed, tonemap_timed;
  char err[512];
};

namespace {

int fail(buddha_ctx *ctx, int code, const char *fmt, ...) {
  char *dst = ctx ? ctx->err : g_create_error;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(ctx, call)                                                                       \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail((ctx), BUDDHA_ECUDA, "CUDA error %d (%s) in %s, line %d (%s)", (int)e_,   \
                  cudaGetErrorString(e_), __FILE__, __LINE__, #call);                       \
  } while (0)

// Constants of the division-free binning for one axis (see orbit_bin and DESIGN.md section 4).
// T = fma(X, inv_half, c0) = (X/2 - min) * inv + 1.5*2^40 +- 2^-11, rounded onto the 2^-12 grid.
FastBin make_fast_bin(double min_v, double delta, int n) {
  FastBin f;
  f.ok = false;
  f.inv_half = 0.0;
  f.c0 = nan("");                                    // (no point is "in range": every one takes bin_exact)
  double inv = 1.0 / delta;
  if (!(inv > 0.0) || !isfinite(inv)) return f;
  if (n > (1 << 19)) return f;                       // quotient must stay below 2^20
  if (!(fabs(min_v) * inv < 0x1p36)) return f;       // keeps |c0| in the 2^40 binade, errors < 2^-12
  long double base = (long double)0x1.8p40 - (long double)min_v * (long double)inv;
  f.c0 = (double)(base + (long double)0x1p-11);     // the upper side: T - 2^-10 < Q < T
  f.inv_half = inv * 0.5;
  f.ok = true;
  return f;
}

void fill_render_params(buddha_ctx *c) {
  const buddha_params &p = c->params;
  RenderParams &r = c->rp;
  memset(&r, 0, sizeof(r));
  r.w = p.width; r.h = p.height;
  r.min_re = p.min_real; r.min_im = p.min_imag;
  r.delta_re = c->delta_re; r.delta_im = c->delta_im;
  FastBin fr = make_fast_bin(p.min_real, c->delta_re, p.width);
  FastBin fi = make_fast_bin(p.min_imag, c->delta_im, p.height);
  r.fast_bin = (fr.ok && fi.ok && !(p.flags & BUDDHA_F_EXACT_BINNING)) ? 1 : 0;
  r.inv_half_re = fr.inv_half; r.c0_re = fr.c0;
  r.inv_half_im = fi.inv_half; r.c0_im = fi.c0;
  if (!r.fast_bin) r.c0_re = r.c0_im = nan("");
  r.max_it = p.max_iterations; r.min_it = p.min_iterations;
  r.shortcut = (p.flags & BUDDHA_F_NO_SHORTCUT) ? 0 : 1;
  r.ship = (p.flags & BUDDHA_F_BURNING_SHIP) ? 1 : 0;
  r.n_ch = c->n_ch > 1 ? c->n_ch : 0;
  r.band_stride = (uint32_t)c->ch_cells;
  if (c->n_ch > 1) {
    // cut the escape-step axis (it = 1-based step of the escape; channel k takes
    // channel_min[k] + 1 <= it <= channel_max[k]) at every window edge; each distinct non-empty
    // set of accepting channels becomes a band
    std::vector<long long> cuts;
    for (int k = 0; k < c->n_ch; k++) {
      cuts.push_back((long long)p.channel_min[k] + 1);
      cuts.push_back((long long)p.channel_max[k] + 1);
    }
    std::sort(cuts.begin(), cuts.end());
    cuts.erase(std::unique(cuts.begin(), cuts.end()), cuts.end());
    c->n_bands = 0;
    r.n_seg = 0;
    for (size_t i = 0; i + 1 < cuts.size(); i++) {
      unsigned set = 0;
      for (int k = 0; k < c->n_ch; k++)
        if (cuts[i] >= (long long)p.channel_min[k] + 1 && cuts[i] <= (long long)p.channel_max[k])
          set |= 1u << k;
      int band = -1;
      if (set) {
        for (int b = 0; b < c->n_bands; b++) if (c->band_set[b] == set) band = b;
        if (band < 0) { band = c->n_bands++; c->band_set[band] = set; }
      }
      const long long lo = std::max<long long>(cuts[i], -0x7fffffffLL);
      r.seg_start[r.n_seg] = (int32_t)std::min<long long>(lo, 0x7fffffffLL);
      r.seg_band[r.n_seg] = band;
      r.n_seg++;
    }
    r.seg_start[r.n_seg] = (int32_t)std::min<long long>(cuts.back(), 0x7fffffffLL);
    r.n_bands = c->n_bands;
  }
  r.ch_low = p.max_iterations;
  for (int k = 0; k < c->n_ch && c->n_ch > 1; k++) {
    r.ch_max[k] = p.channel_max[k];
    r.ch_low = std::min(r.ch_low, p.channel_max[k]);
  }
  uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
  for (int i = 0; i < 10; i++) {
    r.key0[i] = k0; r.key1[i] = k1;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

double now_seconds() {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + (double)ts.tv_nsec / 1e9;
}

// ---- the reference's tone-map expression, evaluated on the host for one count ---------------
// DoGammaCorrection + Clamp (cudabrot.cu:443-449, :416-420).  Same operations in the same order,
// each individually rounded (this TU is compiled with -ffp-contract=off and no FMA ISA), and the
// same libm pow the reference would call on this machine.

inline uint16_t double_to_u16_like_x86(double v) {
  // (uint16_t) of a double on x86-64: cvttsd2si to 32 bits, keep the low half; NaN -> 0.
  if (!(v > -2147483649.0 && v < 2147483648.0)) return 0;
  return (uint16_t)(uint32_t)(int32_t)v;
}

inline uint16_t tone_value(uint32_t count, double scale, double gamma) {
  const double top = 0xffff;
  double scaled = ((double)count) * scale;
  if (gamma <= 0.0) return double_to_u16_like_x86(scaled);
  // pow(x, 1.0) returns x exactly (the result is representable and glibc's pow is within 1 ulp;
  // checked over 2e8 arguments), so the default -g 1.0 needs no libm call per table entry
  const double e = 1 / gamma;
  double v = top * ((e == 1.0) ? (scaled / top) : pow(scaled / top, e));
  if (v != v) return 0;
  if (v <= 0) return 0;
  if (v >= 0xffff) return 0xffff;
  return (uint16_t)v;
}

inline uint16_t bswap16(uint16_t v) { return (uint16_t)((v << 8) | (v >> 8)); }

// ---- NCCL, loaded lazily -----------------------------------------------------------------------

struct NcclApi {
  void *lib;
  int (*CommInitAll)(void **, int, const int *);
  int (*CommDestroy)(void *);
  int (*GroupStart)();
  int (*GroupEnd)();
  int (*Reduce)(const void *, void *, size_t, int, int, int, void *, cudaStream_t);
  const char *(*GetErrorString)(int);
};

NcclApi *load_nccl() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api.lib ? &api : nullptr;
  tried = true;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *n : names) {
    api.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (api.lib) break;
  }
  if (!api.lib) return nullptr;
  api.CommInitAll = (int (*)(void **, int, const int *))dlsym(api.lib, "ncclCommInitAll");
  api.CommDestroy = (int (*)(void *))dlsym(api.lib, "ncclCommDestroy");
  api.GroupStart = (int (*)())dlsym(api.lib, "ncclGroupStart");
  api.GroupEnd = (int (*)())dlsym(api.lib, "ncclGroupEnd");
  api.Reduce = (int (*)(const void *, void *, size_t, int, int, int, void *, cudaStream_t))
      dlsym(api.lib, "ncclReduce");
  api.GetErrorString = (const char *(*)(int))dlsym(api.lib, "ncclGetErrorString");
  if (!api.CommInitAll || !api.CommDestroy || !api.GroupStart || !api.GroupEnd || !api.Reduce) {
    dlclose(api.lib);
    api.lib = nullptr;
    return nullptr;
  }
  return &api;
}

}  // namespace

// ---------------------------------------------------------------------------------------------

extern "C" {

uint32_t buddha_abi_version(void) { return BUDDHA_ABI_VERSION; }

void buddha_default_params(buddha_params *p) {
  memset(p, 0, sizeof(*p));
  p->struct_size = sizeof(*p);
  p->device = 0;
  p->width = 1000; p->height = 1000;
  p->min_real = -2.0; p->max_real = 2.0; p->min_imag = -2.0; p->max_imag = 2.0;
  p->max_iterations = 100; p->min_iterations = 20;
  p->seed = 1337;
}

int buddha_validate_canvas(const buddha_params *p, double *delta_real, double *delta_imag,
                           const char **why) {
  const char *msg = nullptr;
  // same rules, same order, same messages as cudabrot.cu:505-523 (the last one is worded
  // backwards in the reference; kept so scripts that grep the output keep working)
  if (p->width <= 0) msg = "Output width must be positive.";
  else if (p->height <= 0) msg = "Output height must be positive.";
  else if (p->max_real <= p->min_real)
    msg = "Maximum real value must be greater than minimum real value.";
  else if (p->max_imag <= p->min_imag)
    msg = "Minimum imaginary value must be greater than maximum imaginary value.";
  if (why) *why = msg;
  if (msg) return BUDDHA_EINVAL;
  if (delta_imag) *delta_imag = (p->max_imag - p->min_imag) / ((double)p->height);
  if (delta_real) *delta_real = (p->max_real - p->min_real) / ((double)p->width);
  return BUDDHA_OK;
}

const char *buddha_last_error(const buddha_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

int buddha_create(buddha_ctx **out, const buddha_params *p) {
  if (!out || !p) return fail(nullptr, BUDDHA_EINVAL, "null argument");
  *out = nullptr;
  if (p->struct_size != sizeof(buddha_params))
    return fail(nullptr, BUDDHA_EINVAL, "buddha_params.struct_size %u != %zu", p->struct_size,
                sizeof(buddha_params));
  const char *why = nullptr;
  double dre = 0, dim = 0;
  if (buddha_validate_canvas(p, &dre, &dim, &why)) return fail(nullptr, BUDDHA_EINVAL, "%s", why);
  // iteration counts are 32-bit ints as in the reference; the kernels add up to one round on top
  if (p->max_iterations > 0x7fffffff - 4 * kBlock)
    return fail(nullptr, BUDDHA_EINVAL, "max iterations must stay below 2^31 - %d", 4 * kBlock);
  // the reference indexes pixels with 32-bit int (cudabrot.cu:312, :551)
  if ((uint64_t)p->width * (uint64_t)p->height > 0x7fffffffull)
    return fail(nullptr, BUDDHA_EINVAL, "width*height exceeds 2^31-1 cells");

  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(nullptr, BUDDHA_ENODEV, "no CUDA device: %s (libbuddha has no CPU fallback)",
                cudaGetErrorString(e));
  if (p->device < 0 || p->device >= ndev)
    return fail(nullptr, BUDDHA_EINVAL, "device %d out of range (%d devices)", p->device, ndev);
  cudaDeviceProp prop;
  CU(nullptr, cudaGetDeviceProperties(&prop, p->device));
  if (prop.major != 10)
    return fail(nullptr, BUDDHA_ENODEV, "device %d is sm_%d%d; libbuddha is built for sm_100a only",
                p->device, prop.major, prop.minor);
  CU(nullptr, cudaSetDevice(p->device));

  const int n_ch = p->n_channels > 1 ? (int)p->n_channels : 1;
  if (n_ch > BUDDHA_MAX_CHANNELS)
    return fail(nullptr, BUDDHA_EINVAL, "at most %d channels", BUDDHA_MAX_CHANNELS);
  if (n_ch > 1) {
    for (int k = 0; k < n_ch; k++)
      if (p->channel_max[k] <= kT2End || p->channel_max[k] >= (1 << kOrbStepBits))
        return fail(nullptr, BUDDHA_EINVAL, "channel %d: max iterations must be > %d and < 2^%d", k,
                    kT2End, kOrbStepBits);
    if (p->flags & BUDDHA_F_SIMPLE_KERNEL)
      return fail(nullptr, BUDDHA_EINVAL, "the debug kernel renders one channel only");
    for (int k = 0; k < n_ch; k++)
      if (p->channel_min[k] < 0 || p->channel_min[k] >= p->channel_max[k])
        return fail(nullptr, BUDDHA_EINVAL, "channel %d: need 0 <= min < max iterations", k);
    if ((uint64_t)p->width * (uint64_t)p->height * (uint64_t)kMaxBands > 0xffffffffull)
      return fail(nullptr, BUDDHA_EINVAL, "width*height too large for a fused context");
  }

  buddha_ctx *c = (buddha_ctx *)calloc(1, sizeof(buddha_ctx));
  if (!c) return fail(nullptr, BUDDHA_ENOMEM, "out of host memory");
  c->params = *p;
  if (n_ch > 1) {  // the fused pass runs with the widest limits
    c->params.max_iterations = p->channel_max[0];
    c->params.min_iterations = p->channel_min[0];
    for (int k = 1; k < n_ch; k++) {
      c->params.max_iterations = std::max(c->params.max_iterations, p->channel_max[k]);
      c->params.min_iterations = std::min(c->params.min_iterations, p->channel_min[k]);
    }
  }
  c->delta_re = dre; c->delta_im = dim;
  c->n_ch = n_ch;
  c->ch_cells = (size_t)p->width * (size_t)p->height;
  c->cells = c->ch_cells * (size_t)n_ch;
  c->n_bands = 0;
  c->sm_count = prop.multiProcessorCount;
  fill_render_params(c);

  int per_sm = 0;
  // the work stacks are dynamic shared memory (66 KB per CTA: opt-in above 48 KB)
  c->render_smem = kQueueBytes;
  if (const char *e = getenv("BUDDHA_PAD_SMEM")) c->smem_pad = (size_t)atoi(e);  // occupancy experiments
  c->variant = ((p->flags & BUDDHA_F_BURNING_SHIP) ? kVarShip : 0) | (n_ch > 1 ? kVarFused : 0);
  // tiled contexts (decided from the size of the device histogram) take the lean-register build
  {
    const char *e;
    size_t min_mb = 640;
    if ((e = getenv("BUDDHA_TILE_MIN_MB"))) min_mb = (size_t)strtoull(e, nullptr, 10);
    const size_t bands = n_ch > 1 ? (size_t)c->n_bands : 1;
    c->tiled = !(p->flags & BUDDHA_F_SIMPLE_KERNEL) &&
               ((p->flags & BUDDHA_F_FORCE_TILED) != 0 ||
                c->ch_cells * bands * sizeof(uint32_t) >= (min_mb << 20));
  }
  const void *render_fns[8] = {(const void *)render_persistent_kernel<0, kRegsWide>,
                               (const void *)render_persistent_kernel<1, kRegsWide>,
                               (const void *)render_persistent_kernel<2, kRegsWide>,
                               (const void *)render_persistent_kernel<3, kRegsWide>,
                               (const void *)render_persistent_kernel<0, kRegsLean>,
                               (const void *)render_persistent_kernel<1, kRegsLean>,
                               (const void *)render_persistent_kernel<2, kRegsLean>,
                               (const void *)render_persistent_kernel<3, kRegsLean>};
  // The cycle certificate (cert_phase in buddha_kernels.cuh): plain render (not fused, not burning
  // ship) in the 80-register build, shortcuts allowed, and -m large enough for the saved iterations
  // to outweigh the certificate.  A separate instantiation of the kernel.
  size_t cert_cap = 0;
  if (c->variant == 0 && !c->tiled &&
      !(p->flags & (BUDDHA_F_SIMPLE_KERNEL | BUDDHA_F_NO_SHORTCUT)) &&
      p->max_iterations >= kCertMinIt) {
    cert_cap = kCertQueue;
    if (const char *e = getenv("BUDDHA_CERT_QUEUE")) cert_cap = (size_t)strtoull(e, nullptr, 10);  // A/B runs; 0 = off
    if (cert_cap < 2 * kCertHeadroom + 96) cert_cap = 0;
  }
  const void *render_fn = cert_cap ? (const void *)render_persistent_kernel<0, kRegsWide, true>
                                   : render_fns[c->variant + (c->tiled ? 4 : 0)];
  cudaFuncAttributes fattr;
  cudaError_t st = cudaFuncGetAttributes(&fattr, render_fn);
  if (st == cudaSuccess)
    st = cudaFuncSetAttribute(render_fn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              (int)(prop.sharedMemPerBlockOptin - fattr.sharedSizeBytes));
  if (st == cudaSuccess)
    st = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, render_fn, kThreadsPerCta,
                                                       c->render_smem + c->smem_pad);
  if (st != cudaSuccess || per_sm < 1) {
    free(c);
    return fail(nullptr, BUDDHA_ECUDA, "render kernel not launchable on device %d: %s", p->device,
                cudaGetErrorString(st));
  }
  c->grid = per_sm * c->sm_count;

#define CUC(call)                                                                            \
  do {                                                                                       \
    cudaError_t e_ = (call);                                                                 \
    if (e_ != cudaSuccess) {                                                                 \
      fail(nullptr, e_ == cudaErrorMemoryAllocation ? BUDDHA_ENOMEM : BUDDHA_ECUDA,          \
           "CUDA error %d (%s) in %s, line %d (%s)", (int)e_, cudaGetErrorString(e_),        \
           __FILE__, __LINE__, #call);                                                       \
      buddha_destroy(c);                                                                     \
      return e_ == cudaErrorMemoryAllocation ? BUDDHA_ENOMEM : BUDDHA_ECUDA;                 \
    }                                                                                        \
  } while (0)
  CUC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CUC(cudaEventCreate(&c->ev_a));
  CUC(cudaEventCreate(&c->ev_b));
  CUC(cudaEventCreate(&c->ev_ta));
  CUC(cudaEventCreate(&c->ev_tb));
  c->dev_cells = c->ch_cells * (size_t)(n_ch > 1 ? c->n_bands : 1);
  // tiny single-channel canvases get privatised copies (RenderParams::n_copies): up to 64 of
  // them, together at most 4 MB
  c->rp.n_copies = 1;
  c->rp.copy_stride = 0;
  if (n_ch == 1 && !c->tiled && !(p->flags & BUDDHA_F_SIMPLE_KERNEL) &&
      c->dev_cells * sizeof(uint32_t) <= (128u << 10)) {
    c->rp.n_copies = (uint32_t)std::min<size_t>(64, ((size_t)4 << 20) / (c->dev_cells * sizeof(uint32_t)));
    c->rp.copy_stride = (uint32_t)c->dev_cells;
  }
  CUC(cudaMalloc(&c->d_hist, c->dev_cells * c->rp.n_copies * sizeof(uint32_t)));
  CUC(cudaMemsetAsync(c->d_hist, 0, c->dev_cells * c->rp.n_copies * sizeof(uint32_t), c->stream));
  CUC(cudaMalloc(&c->d_cursor, sizeof(unsigned long long)));
  CUC(cudaMalloc(&c->d_counters, sizeof(unsigned long long) * kTotalCnt));
  CUC(cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long) * kTotalCnt, c->stream));
  for (int b = 0; b < 2; b++) {
    c->spill[b].capacity = (unsigned)c->grid * kWarpsPerCta * kSpillPerWarp;
    CUC(cudaMalloc(&c->spill[b].entries, sizeof(double4) * c->spill[b].capacity));
    CUC(cudaMalloc(&c->spill[b].steps, sizeof(int) * c->spill[b].capacity));
    CUC(cudaMalloc(&c->spill[b].count, sizeof(unsigned int) * 4));  // see OrbitSpill
    c->d_spill_next[b] = c->spill[b].count + 1;
  }
  {
    // tile-binned scatter: on for histograms >= 640 MB (config 3: 1.6 GB; crossover measured with tools/gpu_threshold.py), or when forced (tests)
    const char *e;
    const bool forced = (p->flags & BUDDHA_F_FORCE_TILED) != 0;
    if (c->tiled) {
      c->tile_shift = forced ? 12 : 24;  // 16 KB test tiles / 64 MB production tiles
      if ((e = getenv("BUDDHA_TILE_SHIFT"))) c->tile_shift = atoi(e);
      if (c->tile_shift < 8 || c->tile_shift > 28) c->tile_shift = 24;
      while (((c->dev_cells + ((size_t)1 << c->tile_shift) - 1) >> c->tile_shift) > 512 &&
             c->tile_shift < 28)
        c->tile_shift++;  // (the list table in shared memory holds at most 512 tiles)
      c->n_tiles = (int)((c->dev_cells + ((size_t)1 << c->tile_shift) - 1) >> c->tile_shift);
      c->tile_smem = (size_t)c->n_tiles * kWarpsPerCta * sizeof(uint2);
      if (c->n_tiles > 512) {  // not a case tiling is meant for
        buddha_destroy(c);
        return fail(nullptr, BUDDHA_EINVAL, "tile-binned scatter supports at most 512 tiles");
      }
      // the per-warp append counters are dynamic shared memory: keep the grid one resident wave
      int per_sm_tiled = 0;
      CUC(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_tiled, render_fn, kThreadsPerCta,
                                                        c->render_smem + c->smem_pad + c->tile_smem));
      if (per_sm_tiled >= 1) c->grid = per_sm_tiled * c->sm_count;
      c->tile_warps = (uint32_t)c->grid * kWarpsPerCta;
      c->n_lists = (size_t)c->n_tiles * c->tile_warps;
      size_t free_b = 0, total_b = 0;
      CUC(cudaMemGetInfo(&free_b, &total_b));
      size_t pool_b = forced ? ((size_t)4 << 20) : std::min<size_t>((size_t)16 << 30, free_b / 4);
      if ((e = getenv("BUDDHA_TILE_POOL_MB"))) pool_b = (size_t)strtoull(e, nullptr, 10) << 20;
      c->pool_entries = std::min<size_t>(pool_b / sizeof(uint32_t), 0xfffffff0u);  // 32-bit slots
      CUC(cudaMalloc(&c->d_pool, c->pool_entries * sizeof(uint32_t)));
      CUC(cudaMalloc(&c->d_tcount, sizeof(uint32_t) * c->n_lists * 2));
      {
        // high priority: apply CTAs slip into the space the resident render CTAs leave free
        int prio_lo = 0, prio_hi = 0;
        CUC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CUC(cudaStreamCreateWithPriority(&c->apply_stream, cudaStreamNonBlocking, prio_hi));
      }
      CUC(cudaEventCreateWithFlags(&c->ev_rendered, cudaEventDisableTiming));
      CUC(cudaEventCreateWithFlags(&c->ev_applied[0], cudaEventDisableTiming));
      CUC(cudaEventCreateWithFlags(&c->ev_applied[1], cudaEventDisableTiming));
      {
        double4 *ce = nullptr; int *cs = nullptr; unsigned int *cc = nullptr;
        CUC(cudaMalloc(&ce, sizeof(double4) * (size_t)c->tile_warps * 32));
        CUC(cudaMalloc(&cs, sizeof(int) * (size_t)c->tile_warps * 32));
        CUC(cudaMalloc(&cc, sizeof(unsigned int) * c->tile_warps));
        CUC(cudaMemsetAsync(cc, 0, sizeof(unsigned int) * c->tile_warps, c->stream));
        for (int k = 0; k < 2; k++) {
          c->spill[k].carry_entries = ce; c->spill[k].carry_steps = cs; c->spill[k].carry_count = cc;
        }
      }
      CUC(cudaMalloc(&c->d_tcap, sizeof(uint32_t) * c->n_tiles));
      CUC(cudaMalloc(&c->d_tbase, sizeof(uint32_t) * c->n_tiles));
      CUC(cudaMemsetAsync(c->d_tcap, 0, sizeof(uint32_t) * c->n_tiles, c->stream));
      CUC(cudaMemsetAsync(c->d_tbase, 0, sizeof(uint32_t) * c->n_tiles, c->stream));
      c->rp.tile_shift = c->tile_shift;
      c->rp.n_tiles = c->n_tiles;
      c->rp.n_warps = c->tile_warps;
      c->rp.tcount = c->d_tcount; c->rp.tile_cap = c->d_tcap; c->rp.tile_base = c->d_tbase;
      c->rp.pool = c->d_pool;
    }
  }
  if (cert_cap) {   // the certificate queues: cert_cap entries per warp of the render grid
    const size_t entries = (size_t)c->grid * kWarpsPerCta * cert_cap;
    CUC(cudaMalloc(&c->d_cert_c, sizeof(double2) * entries));
    CUC(cudaMalloc(&c->d_cert_z, sizeof(double2) * entries));
    CUC(cudaMalloc(&c->d_cert_m, sizeof(uint4) * entries));
    c->rp.cert_c = c->d_cert_c; c->rp.cert_z = c->d_cert_z; c->rp.cert_m = c->d_cert_m;
    c->rp.cert_cap = (uint32_t)cert_cap;
  }
  CUC(cudaStreamSynchronize(c->stream));
#undef CUC
  *out = c;
  return BUDDHA_OK;
}

void buddha_destroy(buddha_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->params.device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  cudaFree(c->d_hist); cudaFree(c->d_cursor); cudaFree(c->d_counters);
  for (int b = 0; b < 2; b++) {
    cudaFree(c->spill[b].entries); cudaFree(c->spill[b].steps); cudaFree(c->spill[b].count);
  }
  cudaFree(c->spill[0].carry_entries); cudaFree(c->spill[0].carry_steps);
  cudaFree(c->spill[0].carry_count);
  if (c->apply_stream) { cudaStreamSynchronize(c->apply_stream); cudaStreamDestroy(c->apply_stream); }
  if (c->ev_rendered) cudaEventDestroy(c->ev_rendered);
  if (c->ev_applied[0]) cudaEventDestroy(c->ev_applied[0]);
  if (c->ev_applied[1]) cudaEventDestroy(c->ev_applied[1]);
  cudaFree(c->d_pool); cudaFree(c->d_tcount); cudaFree(c->d_tcap); cudaFree(c->d_tbase);
  cudaFree(c->d_cert_c); cudaFree(c->d_cert_z); cudaFree(c->d_cert_m);
  for (int k = 0; k < 2; k++) {
    cudaFree(c->tone[k].d_max); cudaFree(c->tone[k].d_gray); cudaFree(c->tone[k].d_lut);
    cudaFree(c->tone[k].d_thr); cudaFree(c->tone[k].d_chan);
  }
  if (c->h2d_stream) { cudaStreamSynchronize(c->h2d_stream); cudaStreamDestroy(c->h2d_stream); }
  if (c->d2h_stream) { cudaStreamSynchronize(c->d2h_stream); cudaStreamDestroy(c->d2h_stream); }
  if (c->ev_staged) cudaEventDestroy(c->ev_staged);
  if (c->ev_added) cudaEventDestroy(c->ev_added);
  if (c->ev_snap) cudaEventDestroy(c->ev_snap);
  cudaFree(c->d_stage); cudaFree(c->d_snap); cudaFree(c->d_snap_preload); cudaFree(c->d_digest);
  cudaFree(c->d_planes);
  cudaFree(c->d_preload);
  if (c->ev_a) cudaEventDestroy(c->ev_a);
  if (c->ev_b) cudaEventDestroy(c->ev_b);
  if (c->ev_ta) cudaEventDestroy(c->ev_ta);
  if (c->ev_tb) cudaEventDestroy(c->ev_tb);
  if (c->stream) cudaStreamDestroy(c->stream);
  free(c);
}

int buddha_clear_histogram(buddha_ctx *c) {
  if (!c) return BUDDHA_EINVAL;
  CU(c, cudaSetDevice(c->params.device));
  CU(c, cudaMemsetAsync(c->d_hist, 0, c->dev_cells * sizeof(uint32_t), c->stream));
  if (c->d_preload)
    CU(c, cudaMemsetAsync(c->d_preload, 0, c->cells * sizeof(uint32_t), c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return BUDDHA_OK;
}

// Fused contexts keep one histogram per BAND on the device (RenderParams); the host API is
// channel-major: a channel is assembled (sum of its bands + what was loaded) into one buffer.
// `slot` 0 = the live histogram on the render stream, 1 = the snapshot on the copy stream.
static const uint32_t *hist_of(buddha_ctx *c, int slot) { return slot ? c->d_snap : c->d_hist; }
static const uint32_t *preload_of(buddha_ctx *c, int slot) {
  return slot ? c->d_snap_preload : c->d_preload;
}
static cudaStream_t stream_of(buddha_ctx *c, int slot) { return slot ? c->d2h_stream : c->stream; }

// What the host reads as channel `channel`: the histogram itself, or for a fused context the
// channel assembled into the slot's scratch buffer.
static int channel_source(buddha_ctx *c, int slot, int channel, const uint32_t **src) {
  *src = hist_of(c, slot);
  if (c->n_ch < 2) return BUDDHA_OK;
  buddha_ctx::ToneBufs &tb = c->tone[slot];
  if (!tb.d_chan) CU(c, cudaMalloc(&tb.d_chan, sizeof(uint32_t) * c->ch_cells));
  unsigned bands = 0;
  for (int b = 0; b < c->n_bands; b++)
    if ((c->band_set[b] >> channel) & 1u) bands |= 1u << b;
  const uint32_t *pre = preload_of(c, slot);
  channel_sum_kernel<<<c->sm_count * 8, 256, 0, stream_of(c, slot)>>>(
      hist_of(c, slot), pre ? pre + (size_t)channel * c->ch_cells : nullptr, tb.d_chan,
      c->ch_cells, bands);
  CU(c, cudaGetLastError());
  *src = tb.d_chan;
  return BUDDHA_OK;
}

int buddha_load_histogram(buddha_ctx *c, const uint32_t *host, size_t cells) {
  if (!c || !host) return BUDDHA_EINVAL;
  if (cells != c->cells)
    return fail(c, BUDDHA_ESIZE, "histogram has %zu cells, canvas needs %zu", cells, c->cells);
  CU(c, cudaSetDevice(c->params.device));
  if (c->n_ch > 1) {
    // the loaded counts stay channel-major next to the (cleared) bands and are added on read
    if (!c->d_preload) CU(c, cudaMalloc(&c->d_preload, sizeof(uint32_t) * c->cells));
    CU(c, cudaMemsetAsync(c->d_hist, 0, c->dev_cells * sizeof(uint32_t), c->stream));
    CU(c, cudaMemcpyAsync(c->d_preload, host, cells * sizeof(uint32_t), cudaMemcpyHostToDevice,
                          c->stream));
  } else {
    CU(c, cudaMemcpyAsync(c->d_hist, host, cells * sizeof(uint32_t), cudaMemcpyHostToDevice,
                          c->stream));
  }
  CU(c, cudaStreamSynchronize(c->stream));
  return BUDDHA_OK;
}

static int read_channel_impl(buddha_ctx *c, int slot, int channel, uint32_t *host, size_t cells) {
  if (!c || !host) return BUDDHA_EINVAL;
  if (channel < 0 || channel >= c->n_ch) return fail(c, BUDDHA_EINVAL, "no channel %d", channel);
  if (cells != c->ch_cells)
    return fail(c, BUDDHA_ESIZE, "buffer has %zu cells, a channel has %zu", cells, c->ch_cells);
  CU(c, cudaSetDevice(c->params.device));
  const uint32_t *src = nullptr;
  int rc = channel_source(c, slot, channel, &src);
  if (rc) return rc;
  cudaStream_t st = stream_of(c, slot);
  CU(c, cudaMemcpyAsync(host, src, cells * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU(c, cudaStreamSynchronize(st));
  return BUDDHA_OK;
}

int buddha_read_channel(buddha_ctx *c, int channel, uint32_t *host, size_t cells) {
  return read_channel_impl(c, 0, channel, host, cells);
}

int buddha_read_histogram(buddha_ctx *c, uint32_t *host, size_t cells) {
  if (!c || !host) return BUDDHA_EINVAL;
  if (cells != c->cells)
    return fail(c, BUDDHA_ESIZE, "buffer has %zu cells, canvas has %zu", cells, c->cells);
  for (int k = 0; k < c->n_ch; k++) {
    int rc = buddha_read_channel(c, k, host + (size_t)k * c->ch_cells, c->ch_cells);
    if (rc) return rc;
  }
  return BUDDHA_OK;
}

// BUDDHA_TILE_TRACE=1 (experiments): device timestamps of every pipeline launch -- render start /
// end on the main stream, drain end and apply end on the side stream -- printed to stderr.
struct TraceRec { cudaEvent_t r0, r1, d1, a1; };
static std::vector<TraceRec> g_trace;
static bool trace_on() { static const bool on = getenv("BUDDHA_TILE_TRACE") != nullptr; return on; }
static void trace_dump() {
  if (g_trace.empty()) return;
  cudaEvent_t t0 = g_trace[0].r0;
  for (size_t i = 0; i < g_trace.size(); i++) {
    float a = 0, b = 0, d = 0, e = 0;
    cudaEventSynchronize(g_trace[i].a1);
    cudaEventElapsedTime(&a, t0, g_trace[i].r0); cudaEventElapsedTime(&b, t0, g_trace[i].r1);
    cudaEventElapsedTime(&d, t0, g_trace[i].d1); cudaEventElapsedTime(&e, t0, g_trace[i].a1);
    fprintf(stderr, "launch %2zu: render %.2f..%.2f ms (%.2f)  drain done %.2f (+%.2f)  apply done %.2f (+%.2f)\n",
            i, a, b, b - a, d, d - b, e, e - d);
  }
  for (size_t i = 0; i < g_trace.size(); i++) {
    cudaEventDestroy(g_trace[i].r0); cudaEventDestroy(g_trace[i].r1);
    cudaEventDestroy(g_trace[i].d1); cudaEventDestroy(g_trace[i].a1);
  }
  g_trace.clear();
}

// One render launch (+ the orbit drain, + the tile apply when tiling is on) for [first, first+count).
static int launch_render(buddha_ctx *c, uint64_t first, uint64_t count, bool carry_in = false,
                         bool carry_out = false) {
  RenderParams rp = c->rp;
  rp.end = first + count;
  if (c->params.flags & BUDDHA_F_SIMPLE_KERNEL) {
    uint64_t want = (count + 255) / 256;
    int grid = (int)std::min<uint64_t>(want, (uint64_t)c->sm_count * 32);
    render_simple_kernel<<<grid, 256, 0, c->stream>>>(rp, first, c->d_hist, c->d_counters);
  } else {
    // the cursor starts at `first`; warps take rp.chunk indices at a time until it passes rp.end:
    // at least ~8 chunks per warp of the full grid, between kMinChunk and kMaxChunk
    {
      uint64_t per = count / ((uint64_t)c->grid * kWarpsPerCta * 8);
      per = std::min<uint64_t>(std::max<uint64_t>(per, kMinChunk), kMaxChunk);
      rp.chunk = (uint32_t)(per / 32 * 32);
    }
    unsigned long long start = first;
    CU(c, cudaMemcpyAsync(c->d_cursor, &start, sizeof(start), cudaMemcpyHostToDevice, c->stream));
    const int b = c->tile_buf;  // which half of the list pool / which spill list this launch uses
    const bool pipelined = c->tiled && c->tile_calibrated;
    if (c->tiled) {
      if (c->apply_pending[b]) {  // the half is free again once its previous apply has finished
        CU(c, cudaStreamWaitEvent(c->stream, c->ev_applied[b], 0));
        c->apply_pending[b] = false;
      }
      rp.tcount = c->d_tcount + (size_t)b * c->n_lists;
      rp.pool = c->d_pool + (size_t)b * (c->pool_entries / 2);
      CU(c, cudaMemsetAsync(rp.tcount, 0, sizeof(uint32_t) * c->n_lists, c->stream));
    }
    CU(c, cudaMemsetAsync(c->spill[b].count, 0, sizeof(unsigned int) * 4, c->stream));
    uint64_t want = (count + rp.chunk - 1) / rp.chunk;  // warps that can get work at all
    uint64_t ctas = (want + kWarpsPerCta - 1) / kWarpsPerCta;
    int grid = (int)std::min<uint64_t>(ctas, (uint64_t)c->grid);
    if (carry_in || carry_out) grid = c->grid;  // the same warps must exist in both launches
    for (int k = 0; k < 2; k++) {
      c->spill[k].carry_in = carry_in ? 1 : 0;
      c->spill[k].carry_out = carry_out ? 1 : 0;
    }
    const size_t dyn = c->tiled ? c->tile_smem : 0;
    const size_t rsmem = c->render_smem + c->smem_pad + dyn;
#define BUDDHA_LAUNCH_VARIANT(KERNEL, GRID, BLOCK, SMEM, STREAM, ...)                         \
  switch (c->variant) {                                                                     \
    case 0: KERNEL<0><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__); break;                   \
    case 1: KERNEL<1><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__); break;                   \
    case 2: KERNEL<2><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__); break;                   \
    default: KERNEL<3><<<GRID, BLOCK, SMEM, STREAM>>>(__VA_ARGS__); break;                  \
  }
    TraceRec tr = {};
    const bool tracing = c->tiled && c->tile_calibrated && trace_on();
    if (tracing) {
      cudaEventCreate(&tr.r0); cudaEventCreate(&tr.r1); cudaEventCreate(&tr.d1); cudaEventCreate(&tr.a1);
      cudaEventRecord(tr.r0, c->stream);
    }
#define BUDDHA_LAUNCH_RENDER(REGS)                                                          \
  switch (c->variant) {                                                                     \
    case 0: render_persistent_kernel<0, REGS><<<grid, kThreadsPerCta, rsmem, c->stream>>>(  \
        rp, c->d_hist, c->d_cursor, c->d_counters, c->spill[b]); break;                     \
    case 1: render_persistent_kernel<1, REGS><<<grid, kThreadsPerCta, rsmem, c->stream>>>(  \
        rp, c->d_hist, c->d_cursor, c->d_counters, c->spill[b]); break;                     \
    case 2: render_persistent_kernel<2, REGS><<<grid, kThreadsPerCta, rsmem, c->stream>>>(  \
        rp, c->d_hist, c->d_cursor, c->d_counters, c->spill[b]); break;                     \
    default: render_persistent_kernel<3, REGS><<<grid, kThreadsPerCta, rsmem, c->stream>>>( \
        rp, c->d_hist, c->d_cursor, c->d_counters, c->spill[b]); break;                     \
  }
    if (rp.cert_cap != 0u) {   // plain render, 80-register build, with the cycle certificate
      render_persistent_kernel<0, kRegsWide, true><<<grid, kThreadsPerCta, rsmem, c->stream>>>(
          rp, c->d_hist, c->d_cursor, c->d_counters, c->spill[b]);
    } else if (c->tiled) { BUDDHA_LAUNCH_RENDER(kRegsLean) } else { BUDDHA_LAUNCH_RENDER(kRegsWide) }
    CU(c, cudaGetLastError());
    if (tracing) cudaEventRecord(tr.r1, c->stream);
    // In a pipeline of launches (tiling) the rest runs on the second stream, next to the render
    // kernel of the following launch.
    static const bool serial = getenv("BUDDHA_TILE_SERIAL") != nullptr;  // experiment switch
    cudaStream_t side = (pipelined && !serial) ? c->apply_stream : c->stream;
    if (pipelined) {
      CU(c, cudaEventRecord(c->ev_rendered, c->stream));
      CU(c, cudaStreamWaitEvent(side, c->ev_rendered, 0));
    }
    // orbits the warps could not run with enough lanes: finished with grid-wide refill
    const int dgrid = (grid * kWarpsPerCta + kDrainWarps - 1) / kDrainWarps;
    const size_t ddyn = c->tiled ? (size_t)c->n_tiles * kDrainWarps * sizeof(uint2) : 0;
    // long leftovers first, one warp per orbit (their z-chain is the critical path of the launch),
    // then the rest with the lane-group size the volume allows; nothing to drain when the
    // leftovers were carried over to the next launch
    if (!carry_out) {
      BUDDHA_LAUNCH_VARIANT(orbit_drain_kernel, dgrid, kDrainWarps * 32, ddyn, side, rp, c->d_hist,
                            c->d_counters, c->spill[b], c->d_spill_next[b], 1);
      BUDDHA_LAUNCH_VARIANT(orbit_drain_kernel, dgrid, kDrainWarps * 32, ddyn, side, rp, c->d_hist,
                            c->d_counters, c->spill[b], c->d_spill_next[b] + 1, 0);
      c->launches += 2;
    }
    if (tracing) cudaEventRecord(tr.d1, side);
    if (pipelined) {
      // apply this launch's lists while the next launch renders
      CU(c, cudaGetLastError());
      const int agrid = (int)((c->tile_warps + kApplyWarps - 1) / kApplyWarps);
      for (int t = 0; t < c->n_tiles; t++)
        apply_tile_kernel<<<agrid, kApplyWarps * 32, 0, side>>>(
            c->d_hist, rp.tcount, c->d_tcap, c->d_tbase, rp.pool, t, c->tile_warps, c->tile_shift);
      CU(c, cudaEventRecord(c->ev_applied[b], side));
      if (tracing) { cudaEventRecord(tr.a1, side); g_trace.push_back(tr); }
      c->apply_pending[b] = true;
      c->tile_buf = b ^ 1;
      c->launches += c->n_tiles;
    }
  }
  CU(c, cudaGetLastError());
  c->candidates += count;
  c->launches += 1;
  return BUDDHA_OK;
}

// Tiling needs to know how the orbit points spread over the tiles before it can split the pool
// into lists: the first (up to) 2^22 samples of the first render call run with all capacities at
// zero -- every increment takes the direct reduction, the list counters still count -- and the
// observed shares size the lists (85 % by share, 15 % spread evenly).  Happens once per context.
static int calibrate_tiles(buddha_ctx *c, uint64_t first, uint64_t count) {
  int rc = launch_render(c, first, count);
  if (rc) return rc;
  std::vector<uint32_t> cnt(c->n_lists);
  CU(c, cudaMemcpyAsync(cnt.data(), c->d_tcount, sizeof(uint32_t) * c->n_lists,
                        cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  // share of every tile (the warps take samples from one shared cursor, so within a tile the
  // lists fill evenly and get equal capacity)
  std::vector<double> tile_pts(c->n_tiles, 0.0);
  double total = 0;
  for (int t = 0; t < c->n_tiles; t++) {
    for (uint32_t w = 0; w < c->tile_warps; w++) tile_pts[t] += cnt[(size_t)t * c->tile_warps + w];
    total += tile_pts[t];
  }
  std::vector<uint32_t> cap(c->n_tiles), base(c->n_tiles);
  uint64_t pos = 0;
  for (int t = 0; t < c->n_tiles; t++) {
    double share = total > 0 ? tile_pts[t] / total : 1.0 / c->n_tiles;
    double want = (double)(c->pool_entries / 2) * (0.85 * share + 0.15 / c->n_tiles) / c->tile_warps;
    cap[t] = (uint32_t)want;  // the shares sum to 1: the lists fit one pool half (< 2^32 entries)
    base[t] = (uint32_t)pos;
    pos += (uint64_t)cap[t] * c->tile_warps;
  }
  CU(c, cudaMemcpyAsync(c->d_tcap, cap.data(), sizeof(uint32_t) * c->n_tiles,
                        cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemcpyAsync(c->d_tbase, base.data(), sizeof(uint32_t) * c->n_tiles,
                        cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->tile_pts_per_sample = total / (double)count;
  c->tile_calibrated = true;
  return BUDDHA_OK;
}

static int enqueue_render(buddha_ctx *c, uint64_t first, uint64_t count) {
  if (count == 0) return BUDDHA_OK;
  // the device cursor overshoots the end of a range by up to one chunk per warp of the grid
  const uint64_t slack = (uint64_t)c->grid * kWarpsPerCta * kMaxChunk;
  if (first + count < first || first + count > ~(uint64_t)0 - slack)
    return fail(c, BUDDHA_EINVAL, "sample range must end below 2^64 - %llu",
                (unsigned long long)slack);
  if (!c->tiled) {
    int rc = launch_render(c, first, count);
    if (rc == BUDDHA_OK && c->rp.n_copies > 1) {
      fold_copies_kernel<<<c->sm_count, 256, 0, c->stream>>>(c->d_hist, c->dev_cells, c->rp.n_copies,
                                                             c->rp.copy_stride);
      CU(c, cudaGetLastError());
      c->launches += 1;
    }
    return rc;
  }
  if (!c->tile_calibrated) {
    uint64_t n0 = std::min<uint64_t>(count, (uint64_t)1 << 22);
    int rc = calibrate_tiles(c, first, n0);
    if (rc) return rc;
    first += n0;
    count -= n0;
  }
  // launches sized so that the expected number of increments fills ~70 % of one pool half, and
  // at most 2^30 samples so that a large call becomes a pipeline of render / apply pairs
  double per = 0.7 * (double)(c->pool_entries / 2) / std::max(c->tile_pts_per_sample, 1e-6);
  double max_launch = 1073741824.0;
  if (const char *e = getenv("BUDDHA_TILE_LAUNCH_LOG2")) max_launch = ldexp(1.0, atoi(e));
  uint64_t max_n = (uint64_t)std::min(std::max(per, 65536.0), max_launch);
  max_n = std::max<uint64_t>(max_n / kChunk * kChunk, kChunk);
  bool carried = false;
  while (count > 0) {
    uint64_t n = std::min(count, max_n);
    static const bool no_carry = getenv("BUDDHA_NO_CARRY") != nullptr;  // experiment switch
    const bool more = count > n && !no_carry;  // leftovers ride along to the next launch of this call
    int rc = launch_render(c, first, n, carried, more);
    if (rc) return rc;
    carried = more;
    first += n;
    count -= n;
  }
  for (int b = 0; b < 2; b++) {  // everything after this call on the main stream sees the result
    if (c->apply_pending[b]) {
      CU(c, cudaStreamWaitEvent(c->stream, c->ev_applied[b], 0));
      c->apply_pending[b] = false;
    }
  }
  return BUDDHA_OK;
}

int buddha_render_samples_async(buddha_ctx *c, uint64_t first, uint64_t count) {
  if (!c) return BUDDHA_EINVAL;
  CU(c, cudaSetDevice(c->params.device));
  CU(c, cudaEventRecord(c->ev_a, c->stream));
  int rc = enqueue_render(c, first, count);
  if (rc) return rc;
  CU(c, cudaEventRecord(c->ev_b, c->stream));
  c->render_timed = true;
  return BUDDHA_OK;
}

int buddha_sync(buddha_ctx *c) {
  if (!c) return BUDDHA_EINVAL;
  CU(c, cudaSetDevice(c->params.device));
  CU(c, cudaStreamSynchronize(c->stream));
  if (trace_on()) trace_dump();
  return BUDDHA_OK;
}

int buddha_render_samples(buddha_ctx *c, uint64_t first, uint64_t count) {
  int rc = buddha_render_samples_async(c, first, count);
  if (rc) return rc;
  return buddha_sync(c);
}

int buddha_render_seconds(buddha_ctx *c, double seconds, volatile int *stop, uint64_t first,
                          uint64_t *samples_done, uint64_t *passes) {
  if (!c) return BUDDHA_EINVAL;
  if (seconds < 0 && !stop)
    return fail(c, BUDDHA_EINVAL, "an endless run (seconds < 0) needs a stop flag");
  CU(c, cudaSetDevice(c->params.device));
  uint64_t done = 0, npass = 0;
  // first pass = one reference pass (512*512*50 candidates, cudabrot.cu:20,23,34); later passes
  // are sized for ~0.1 s so the stop flag / deadline latency stays bounded at any -m
  uint64_t pass = 13107200ull;
  double t0 = now_seconds();
  CU(c, cudaEventRecord(c->ev_a, c->stream));
  while (!(stop && *stop)) {
    double ta = now_seconds();
    int rc = enqueue_render(c, first + done, pass);
    if (rc) return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    double tb = now_seconds();
    done += pass;
    npass++;
    if ((seconds >= 0) && ((tb - t0) > seconds)) break;
    double rate = (double)pass / std::max(tb - ta, 1e-6);
    double next = rate * 0.1;
    if (seconds >= 0) next = std::min(next, rate * std::max(seconds - (tb - t0), 0.005));
    uint64_t n = (uint64_t)std::min(std::max(next, 1048576.0), 68719476736.0);
    pass = (n + kChunk - 1) / kChunk * kChunk;
  }
  CU(c, cudaEventRecord(c->ev_b, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->render_timed = true;
  if (samples_done) *samples_done = done;
  if (passes) *passes = npass;
  return BUDDHA_OK;
}

int buddha_last_render_ms(buddha_ctx *c, float *ms) {
  if (!c || !ms) return BUDDHA_EINVAL;
  if (!c->render_timed) return fail(c, BUDDHA_EINVAL, "no render call has been timed yet");
  CU(c, cudaSetDevice(c->params.device));
  CU(c, cudaEventSynchronize(c->ev_b));
  CU(c, cudaEventElapsedTime(ms, c->ev_a, c->ev_b));
  return BUDDHA_OK;
}

static int read_counters(buddha_ctx *c, unsigned long long *v) {
  CU(c, cudaSetDevice(c->params.device));
  CU(c, cudaMemcpyAsync(v, c->d_counters, sizeof(unsigned long long) * kTotalCnt,
                        cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return BUDDHA_OK;
}

int buddha_get_channel_counters(buddha_ctx *c, int channel, buddha_counters *out) {
  if (!c || !out) return BUDDHA_EINVAL;
  if (c->n_ch < 2) return channel == 0 ? buddha_get_counters(c, out) : BUDDHA_EINVAL;
  if (channel < 0 || channel >= c->n_ch) return fail(c, BUDDHA_EINVAL, "no channel %d", channel);
  unsigned long long v[kTotalCnt];
  int rc = read_counters(c, v);
  if (rc) return rc;
  const unsigned long long *ch = v + kCntSlots + channel * kChSlots;  // kChHit / kChOver: by channel
  unsigned long long accepted = 0, points = 0, increments = 0;       // the rest: by band
  for (int b = 0; b < c->n_bands; b++) {
    if (!((c->band_set[b] >> channel) & 1u)) continue;
    const unsigned long long *bd = v + kCntSlots + b * kChSlots;
    accepted += bd[kChAccepted]; points += bd[kChPoints]; increments += bd[kChIncrements];
  }
  memset(out, 0, sizeof(*out));
  out->candidates = c->candidates;
  out->rejected = v[kCntRejected];
  // the device counts, per channel, the samples that ESCAPED beyond the channel's limit; samples
  // that never escaped (v[kCntHitMax]) hit every channel's limit
  const unsigned long long widest = (unsigned long long)c->params.max_iterations;
  const unsigned long long limit = (unsigned long long)c->params.channel_max[channel];
  out->hit_max = ch[kChHit] + v[kCntHitMax];
  out->accepted = accepted;
  out->too_early = c->candidates - v[kCntRejected] - out->hit_max - accepted;
  // sum over samples of min(steps run, this channel's limit)
  out->escape_iters = v[kCntEscapeIters] - ch[kChOver] - v[kCntHitMax] * (widest - limit);
  out->orbit_points = points;
  out->increments = increments;
  out->executed_iters = v[kCntExecuted];
  out->shortcut_hits = v[kCntShortcut];
  out->exact_bins = v[kCntExactBins];
  out->kernel_launches = c->launches;
  return BUDDHA_OK;
}

int buddha_get_counters(buddha_ctx *c, buddha_counters *out) {
  if (!c || !out) return BUDDHA_EINVAL;
  unsigned long long v[kTotalCnt];
  int rc = read_counters(c, v);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  out->candidates = c->candidates;
  out->rejected = v[kCntRejected];
  out->hit_max = v[kCntHitMax];
  out->accepted = v[kCntAccepted];
  // every candidate ends in exactly one class; too_early is the remainder
  out->too_early = c->candidates - v[kCntRejected] - v[kCntHitMax] - v[kCntAccepted];
  out->escape_iters = v[kCntEscapeIters];
  out->orbit_points = v[kCntOrbitPoints];
  out->increments = v[kCntIncrements];
  out->executed_iters = v[kCntExecuted];
  out->shortcut_hits = v[kCntShortcut];
  out->exact_bins = v[kCntExactBins];
  out->kernel_launches = c->launches;
  return BUDDHA_OK;
}

int buddha_reset_counters(buddha_ctx *c) {
  if (!c) return BUDDHA_EINVAL;
  CU(c, cudaSetDevice(c->params.device));
  CU(c, cudaMemsetAsync(c->d_counters, 0, sizeof(unsigned long long) * kTotalCnt, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  c->candidates = 0;
  c->launches = 0;
  return BUDDHA_OK;
}

int buddha_tonemap_u16(buddha_ctx *c, double gamma, int big_endian, uint16_t *host_out,
                       size_t cells, uint32_t *max_out, double *scale_out) {
  return buddha_tonemap_channel_u16(c, 0, gamma, big_endian, host_out, cells, max_out, scale_out);
}

static int tonemap_impl(buddha_ctx *c, int slot, int channel, double gamma, int big_endian,
                        uint16_t *host_out, size_t cells, uint32_t *max_out, double *scale_out,
                        uint16_t *dev_out = nullptr) {
  // dev_out: leave the image in this device buffer instead of copying it to the host
  if (!c) return BUDDHA_EINVAL;
  if (channel < 0 || channel >= c->n_ch) return fail(c, BUDDHA_EINVAL, "no channel %d", channel);
  if (host_out && cells != c->ch_cells)
    return fail(c, BUDDHA_ESIZE, "image buffer has %zu cells, canvas has %zu", cells, c->ch_cells);
  CU(c, cudaSetDevice(c->params.device));  // before anything allocates or launches
  buddha_ctx::ToneBufs &tb = c->tone[slot];
  cudaStream_t st = stream_of(c, slot);
  const uint32_t *d_src = nullptr;
  int rc = channel_source(c, slot, channel, &d_src);
  if (rc) return rc;
  if (!tb.d_max) CU(c, cudaMalloc(&tb.d_max, sizeof(uint32_t)));
  const int blocks = c->sm_count * 8;

  // pass 1: GetLinearColorScale's maximum (cudabrot.cu:430-435)
  if (slot == 0) CU(c, cudaEventRecord(c->ev_ta, st));
  CU(c, cudaMemsetAsync(tb.d_max, 0, sizeof(uint32_t), st));
  hist_max_kernel<<<blocks, 256, 0, st>>>(d_src, c->ch_cells, tb.d_max);
  CU(c, cudaGetLastError());
  c->launches += 1;
  uint32_t mx = 0;
  CU(c, cudaMemcpyAsync(&mx, tb.d_max, sizeof(mx), cudaMemcpyDeviceToHost, st));
  CU(c, cudaStreamSynchronize(st));
  double scale = ((double)0xffff) / ((double)mx);  // :436 (inf when the histogram is empty)
  if (max_out) *max_out = mx;
  if (scale_out) *scale_out = scale;
  if (!host_out && !dev_out) {
    if (slot == 0) {
      CU(c, cudaEventRecord(c->ev_tb, st));
      c->tonemap_timed = true;
    }
    return BUDDHA_OK;
  }

  // host: tabulate the reference's count -> grey expression (glibc pow and all)
  uint32_t lut_size = (mx < kLutMax) ? (mx + 1) : kLutMax;
  std::vector<uint16_t> lut(lut_size);
#pragma omp parallel for schedule(static)
  for (long long k = 0; k < (long long)lut_size; k++) {
    uint16_t v = tone_value((uint32_t)k, scale, gamma);
    lut[k] = big_endian ? bswap16(v) : v;
  }
  if (lut_size > tb.lut_capacity) {
    cudaFree(tb.d_lut);
    tb.d_lut = nullptr;
    tb.lut_capacity = 0;
    CU(c, cudaMalloc(&tb.d_lut, sizeof(uint16_t) * (size_t)lut_size));
    tb.lut_capacity = lut_size;
  }
  CU(c, cudaMemcpyAsync(tb.d_lut, lut.data(), sizeof(uint16_t) * (size_t)lut_size,
                        cudaMemcpyHostToDevice, st));
  if (!tb.d_thr) CU(c, cudaMalloc(&tb.d_thr, sizeof(uint32_t) * 65536));
  std::vector<uint32_t> thr;
  if (mx >= lut_size) {
    // counts past the table: thr[v] = smallest count in [0, mx] whose value is >= v (mx + 1 if
    // no count reaches v), on the (monotone) reference expression.  A closed-form inverse gives
    // a guess that a short walk makes exact; bisection if the walk does not settle.
    thr.resize(65536);
    const double top = 0xffff;
#pragma omp parallel for schedule(static)
    for (int v = 0; v < 65536; v++) {
      const uint64_t end = (uint64_t)mx + 1;
      double x = (gamma > 0.0) ? top * pow((double)v / top, gamma) / scale : (double)v / scale;
      uint64_t cnt = (x >= 0.0 && x < (double)end) ? (uint64_t)x : end;
      int steps = 0;
      while (cnt > 0 && steps < 64 && tone_value((uint32_t)(cnt - 1), scale, gamma) >= (uint16_t)v) { cnt--; steps++; }
      while (cnt < end && steps < 64 && tone_value((uint32_t)cnt, scale, gamma) < (uint16_t)v) { cnt++; steps++; }
      if (steps >= 64) {
        uint64_t lo = 0, hi = end;
        while (lo < hi) {
          uint64_t mid = (lo + hi) >> 1;
          if (tone_value((uint32_t)mid, scale, gamma) >= (uint16_t)v) hi = mid; else lo = mid + 1;
        }
        cnt = lo;
      }
      thr[v] = (uint32_t)std::min<uint64_t>(cnt, 0xffffffffull);
    }
    thr[0] = 0;
    CU(c, cudaMemcpyAsync(tb.d_thr, thr.data(), sizeof(uint32_t) * 65536, cudaMemcpyHostToDevice,
                          st));
  }
  if (!dev_out && !tb.d_gray) CU(c, cudaMalloc(&tb.d_gray, sizeof(uint16_t) * c->ch_cells));
  uint16_t *d_img = dev_out ? dev_out : tb.d_gray;

  // pass 2: the map itself, 4 B read + 2 B written per pixel
  tonemap_kernel<<<blocks, 256, 0, st>>>(d_src, d_img, c->ch_cells, tb.d_lut, lut_size,
                                         tb.d_thr, big_endian ? 1 : 0);
  CU(c, cudaGetLastError());
  c->launches += 1;
  if (slot == 0) {
    CU(c, cudaEventRecord(c->ev_tb, st));
    c->tonemap_timed = true;
  }
  if (host_out)
    CU(c, cudaMemcpyAsync(host_out, d_img, sizeof(uint16_t) * c->ch_cells, cudaMemcpyDeviceToHost,
                          st));
  CU(c, cudaStreamSynchronize(st));  // (lut / thr are host vectors read by the copies above)
  return BUDDHA_OK;
}

int buddha_tonemap_channel_u16(buddha_ctx *c, int channel, double gamma, int big_endian,
                               uint16_t *host_out, size_t cells, uint32_t *max_out,
                               double *scale_out) {
  return tonemap_impl(c, 0, channel, gamma, big_endian, host_out, cells, max_out, scale_out);
}

int buddha_combine_rgb_u16(buddha_ctx *c, const int channels[3], double gamma, int mode,
                           double hue_adjust, int big_endian, uint16_t *host_rgb, size_t pixels,
                           uint32_t max_out[3]) {
  if (!c || !channels || !host_rgb) return BUDDHA_EINVAL;
  if (pixels != c->ch_cells)
    return fail(c, BUDDHA_ESIZE, "image buffer has %zu pixels, canvas has %zu", pixels, c->ch_cells);
  if (mode != BUDDHA_COMBINE_RGB && mode != BUDDHA_COMBINE_HSL)
    return fail(c, BUDDHA_EINVAL, "unknown combine mode %d", mode);
  CU(c, cudaSetDevice(c->params.device));
  if (!c->d_planes) CU(c, cudaMalloc(&c->d_planes, sizeof(uint16_t) * c->ch_cells * 6));
  for (int k = 0; k < 3; k++) {
    uint32_t mx = 0;
    int rc = tonemap_impl(c, 0, channels[k], gamma, 0, nullptr, 0, &mx, nullptr,
                          c->d_planes + (size_t)k * c->ch_cells);
    if (rc) return rc;
    if (max_out) max_out[k] = mx;
  }
  uint16_t *d_rgb = c->d_planes + 3 * c->ch_cells;
  combine_rgb_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(
      c->d_planes, c->d_planes + c->ch_cells, c->d_planes + 2 * c->ch_cells, d_rgb, c->ch_cells, mode,
      hue_adjust, big_endian ? 1 : 0);
  CU(c, cudaGetLastError());
  c->launches += 1;
  CU(c, cudaMemcpyAsync(host_rgb, d_rgb, sizeof(uint16_t) * 3 * c->ch_cells, cudaMemcpyDeviceToHost,
                        c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return BUDDHA_OK;
}

// ---- asynchronous host transfers ------------------------------------------------------------

static int transfer_setup(buddha_ctx *c) {
  if (c->d2h_stream) return BUDDHA_OK;
  CU(c, cudaStreamCreateWithFlags(&c->h2d_stream, cudaStreamNonBlocking));
  CU(c, cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
  CU(c, cudaEventCreateWithFlags(&c->ev_staged, cudaEventDisableTiming));
  CU(c, cudaEventCreateWithFlags(&c->ev_added, cudaEventDisableTiming));
  CU(c, cudaEventCreateWithFlags(&c->ev_snap, cudaEventDisableTiming));
  return BUDDHA_OK;
}

int buddha_add_histogram_async(buddha_ctx *c, const uint32_t *host, size_t cells) {
  if (!c || !host) return BUDDHA_EINVAL;
  if (cells != c->cells)
    return fail(c, BUDDHA_ESIZE, "histogram has %zu cells, canvas needs %zu", cells, c->cells);
  CU(c, cudaSetDevice(c->params.device));
  int rc = transfer_setup(c);
  if (rc) return rc;
  if (!c->d_stage) CU(c, cudaMalloc(&c->d_stage, sizeof(uint32_t) * c->cells));
  uint32_t *dst = c->d_hist;
  if (c->n_ch > 1) {  // loaded counts stay channel-major beside the bands (buddha_load_histogram)
    if (!c->d_preload) {
      CU(c, cudaMalloc(&c->d_preload, sizeof(uint32_t) * c->cells));
      CU(c, cudaMemsetAsync(c->d_preload, 0, sizeof(uint32_t) * c->cells, c->stream));
    }
    dst = c->d_preload;
  }
  // the staging buffer is free again once the previous add kernel has run; the copy itself only
  // waits for that, not for renders enqueued since, so it overlaps them
  CU(c, cudaStreamWaitEvent(c->h2d_stream, c->ev_added, 0));  // (no-op before the first add)
  CU(c, cudaMemcpyAsync(c->d_stage, host, sizeof(uint32_t) * cells, cudaMemcpyHostToDevice,
                        c->h2d_stream));
  CU(c, cudaEventRecord(c->ev_staged, c->h2d_stream));
  CU(c, cudaStreamWaitEvent(c->stream, c->ev_staged, 0));
  add_cells_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(dst, c->d_stage, cells);
  CU(c, cudaGetLastError());
  CU(c, cudaEventRecord(c->ev_added, c->stream));
  c->launches += 1;
  return BUDDHA_OK;
}

int buddha_snapshot(buddha_ctx *c) {
  if (!c) return BUDDHA_EINVAL;
  CU(c, cudaSetDevice(c->params.device));
  int rc = transfer_setup(c);
  if (rc) return rc;
  if (!c->d_snap) CU(c, cudaMalloc(&c->d_snap, sizeof(uint32_t) * c->dev_cells));
  CU(c, cudaMemcpyAsync(c->d_snap, c->d_hist, sizeof(uint32_t) * c->dev_cells,
                        cudaMemcpyDeviceToDevice, c->stream));
  if (c->d_preload) {
    if (!c->d_snap_preload) CU(c, cudaMalloc(&c->d_snap_preload, sizeof(uint32_t) * c->cells));
    CU(c, cudaMemcpyAsync(c->d_snap_preload, c->d_preload, sizeof(uint32_t) * c->cells,
                          cudaMemcpyDeviceToDevice, c->stream));
  }
  CU(c, cudaEventRecord(c->ev_snap, c->stream));
  CU(c, cudaStreamWaitEvent(c->d2h_stream, c->ev_snap, 0));
  c->snap_valid = true;
  return BUDDHA_OK;
}

int buddha_read_snapshot(buddha_ctx *c, uint32_t *host, size_t cells) {
  if (!c || !host) return BUDDHA_EINVAL;
  if (!c->snap_valid) return fail(c, BUDDHA_EINVAL, "no snapshot has been taken");
  if (cells != c->cells)
    return fail(c, BUDDHA_ESIZE, "buffer has %zu cells, canvas has %zu", cells, c->cells);
  for (int k = 0; k < c->n_ch; k++) {
    int rc = read_channel_impl(c, 1, k, host + (size_t)k * c->ch_cells, c->ch_cells);
    if (rc) return rc;
  }
  return BUDDHA_OK;
}

int buddha_tonemap_snapshot_u16(buddha_ctx *c, int channel, double gamma, int big_endian,
                                uint16_t *host_out, size_t cells, uint32_t *max_out,
                                double *scale_out) {
  if (!c) return BUDDHA_EINVAL;
  if (!c->snap_valid) return fail(c, BUDDHA_EINVAL, "no snapshot has been taken");
  return tonemap_impl(c, 1, channel, gamma, big_endian, host_out, cells, max_out, scale_out);
}

int buddha_histogram_digest(buddha_ctx *c, int channel, uint64_t *digest) {
  if (!c || !digest) return BUDDHA_EINVAL;
  if (channel < 0 || channel >= c->n_ch) return fail(c, BUDDHA_EINVAL, "no channel %d", channel);
  CU(c, cudaSetDevice(c->params.device));
  const uint32_t *src = nullptr;
  int rc = channel_source(c, 0, channel, &src);
  if (rc) return rc;
  const size_t n_blocks = (c->ch_cells + kDigestBlock - 1) / kDigestBlock;
  if (n_blocks > c->digest_capacity) {
    cudaFree(c->d_digest);
    c->d_digest = nullptr;
    c->digest_capacity = 0;
    CU(c, cudaMalloc(&c->d_digest, sizeof(unsigned long long) * n_blocks));
    c->digest_capacity = n_blocks;
  }
  digest_blocks_kernel<<<c->sm_count * 8, 256, 0, c->stream>>>(src, c->ch_cells, c->d_digest,
                                                                n_blocks);
  CU(c, cudaGetLastError());
  c->launches += 1;
  std::vector<unsigned long long> blocks(n_blocks);
  CU(c, cudaMemcpyAsync(blocks.data(), c->d_digest, sizeof(unsigned long long) * n_blocks,
                        cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  unsigned long long h = kFnvBasis;
  for (size_t b = 0; b < n_blocks; b++) h = (h ^ blocks[b]) * kFnvPrime;
  *digest = h;
  return BUDDHA_OK;
}

int buddha_last_tonemap_ms(buddha_ctx *c, float *ms) {
  if (!c || !ms) return BUDDHA_EINVAL;
  if (!c->tonemap_timed) return fail(c, BUDDHA_EINVAL, "no tone-map call has been timed yet");
  CU(c, cudaSetDevice(c->params.device));
  CU(c, cudaEventSynchronize(c->ev_tb));
  CU(c, cudaEventElapsedTime(ms, c->ev_ta, c->ev_tb));
  return BUDDHA_OK;
}

void *buddha_device_histogram(buddha_ctx *c) { return c ? (void *)c->d_hist : nullptr; }
size_t buddha_device_histogram_cells(buddha_ctx *c) { return c ? c->dev_cells : 0; }
void *buddha_stream(buddha_ctx *c) { return c ? (void *)c->stream : nullptr; }

int buddha_merge(buddha_ctx **ctxs, int n, int root) {
  if (!ctxs || n < 1 || root < 0 || root >= n) return BUDDHA_EINVAL;
  if (n == 1) return BUDDHA_OK;
  buddha_ctx *r = ctxs[root];
  if (!r) return BUDDHA_EINVAL;
  bool any_preload = false;
  for (int i = 0; i < n; i++) {
    const buddha_ctx *x = ctxs[i];
    if (!x) return BUDDHA_EINVAL;
    // the cells only add up if every context bins the same canvas into the same bands
    const buddha_params &a = x->params, &b = r->params;
    bool same = a.width == b.width && a.height == b.height && a.min_real == b.min_real &&
                a.max_real == b.max_real && a.min_imag == b.min_imag && a.max_imag == b.max_imag &&
                x->n_ch == r->n_ch && x->n_bands == r->n_bands && x->dev_cells == r->dev_cells &&
                x->rp.n_seg == r->rp.n_seg;
    for (int k = 0; same && k < x->n_bands; k++) same = x->band_set[k] == r->band_set[k];
    for (int k = 0; same && k < x->rp.n_seg; k++)
      same = x->rp.seg_start[k] == r->rp.seg_start[k] && x->rp.seg_band[k] == r->rp.seg_band[k] &&
             x->rp.seg_start[k + 1] == r->rp.seg_start[k + 1];
    same = same && a.max_iterations == b.max_iterations && a.min_iterations == b.min_iterations &&
           ((a.flags ^ b.flags) & BUDDHA_F_BURNING_SHIP) == 0;
    if (!same) return fail(r, BUDDHA_ESIZE, "context %d differs from the root in canvas or channels", i);
    for (int j = 0; j < i; j++)
      if (ctxs[j]->params.device == x->params.device)
        return fail(r, BUDDHA_EINVAL, "contexts %d and %d share device %d", j, i, x->params.device);
    any_preload = any_preload || x->d_preload != nullptr;
  }
  int prev_dev = 0;
  CU(r, cudaGetDevice(&prev_dev));
  // fused contexts keep loaded counts channel-major beside the bands: they are summed as well
  if (any_preload) {
    for (int i = 0; i < n; i++) {
      buddha_ctx *x = ctxs[i];
      if (x->d_preload) continue;
      CU(x, cudaSetDevice(x->params.device));
      CU(x, cudaMalloc(&x->d_preload, sizeof(uint32_t) * x->cells));
      CU(x, cudaMemsetAsync(x->d_preload, 0, sizeof(uint32_t) * x->cells, x->stream));
    }
  }
  NcclApi *nccl = load_nccl();
  if (!nccl) {
    cudaSetDevice(prev_dev);
    return fail(r, BUDDHA_ENCCL, "libnccl.so.2 could not be loaded: %s", dlerror());
  }
  std::vector<int> devs(n);
  std::vector<void *> comms(n, nullptr);
  for (int i = 0; i < n; i++) devs[i] = ctxs[i]->params.device;
  int rc = nccl->CommInitAll(comms.data(), n, devs.data());
  if (rc) {
    cudaSetDevice(prev_dev);
    return fail(r, BUDDHA_ENCCL, "ncclCommInitAll: %s",
                nccl->GetErrorString ? nccl->GetErrorString(rc) : "error");
  }
  rc = nccl->GroupStart();
  for (int i = 0; i < n && !rc; i++) {
    cudaSetDevice(devs[i]);
    rc = nccl->Reduce(ctxs[i]->d_hist, ctxs[i]->d_hist, r->dev_cells, /*ncclUint32*/ 3, /*ncclSum*/ 0,
                      root, comms[i], ctxs[i]->stream);
  }
  int rc2 = nccl->GroupEnd();
  if (!rc) rc = rc2;
  if (!rc && any_preload) {
    rc = nccl->GroupStart();
    for (int i = 0; i < n && !rc; i++) {
      cudaSetDevice(devs[i]);
      rc = nccl->Reduce(ctxs[i]->d_preload, ctxs[i]->d_preload, r->cells, 3, 0, root, comms[i],
                        ctxs[i]->stream);
    }
    rc2 = nccl->GroupEnd();
    if (!rc) rc = rc2;
  }
  for (int i = 0; i < n; i++) {
    cudaSetDevice(devs[i]);
    cudaStreamSynchronize(ctxs[i]->stream);
  }
  for (int i = 0; i < n; i++) nccl->CommDestroy(comms[i]);
  cudaSetDevice(prev_dev);  // the caller's current device is not ours to change
  if (rc) return fail(r, BUDDHA_ENCCL, "ncclReduce: %s",
                      nccl->GetErrorString ? nccl->GetErrorString(rc) : "error");
  return BUDDHA_OK;
}

int buddha_probe_fp64_peak(buddha_ctx *c, double *lane_instr_per_s) {
  if (!c || !lane_instr_per_s) return BUDDHA_EINVAL;
  CU(c, cudaSetDevice(c->params.device));
  double *d_out = nullptr;
  CU(c, cudaMalloc(&d_out, sizeof(double)));
  const int blocks = c->sm_count * 8, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CU(c, cudaEventRecord(c->ev_ta, c->stream));
    probe_dfma_kernel<<<blocks, 256, 0, c->stream>>>(d_out, iters, 1.0000001, 1e-9);
    CU(c, cudaEventRecord(c->ev_tb, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    CU(c, cudaEventElapsedTime(&ms, c->ev_ta, c->ev_tb));
    double rate = (double)blocks * 256.0 * iters * 64.0 / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  cudaFree(d_out);
  c->tonemap_timed = false;
  *lane_instr_per_s = best;
  return BUDDHA_OK;
}

int buddha_probe_red_peak(buddha_ctx *c, size_t footprint_bytes, double *red_per_s) {
  if (!c || !red_per_s || footprint_bytes < 4) return BUDDHA_EINVAL;
  CU(c, cudaSetDevice(c->params.device));
  uint32_t *buf = nullptr;
  unsigned long long cells = footprint_bytes / 4;
  CU(c, cudaMalloc(&buf, cells * 4));
  CU(c, cudaMemsetAsync(buf, 0, cells * 4, c->stream));
  const int blocks = c->sm_count * 8, per_thread = 256;
  double best = 0.0;
  for (int rep = 0; rep < 4; rep++) {
    CU(c, cudaEventRecord(c->ev_ta, c->stream));
    probe_red_kernel<<<blocks, 256, 0, c->stream>>>(buf, cells, per_thread, (uint32_t)rep);
    CU(c, cudaEventRecord(c->ev_tb, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    float ms = 0;
    CU(c, cudaEventElapsedTime(&ms, c->ev_ta, c->ev_tb));
    double rate = (double)blocks * 256.0 * per_thread / (ms * 1e-3);
    if (rep > 0) best = std::max(best, rate);
  }
  cudaFree(buf);
  c->tonemap_timed = false;
  *red_per_s = best;
  return BUDDHA_OK;
}

}  // extern "C"
