"""ctypes binding of libbuddha.so -- one Python function per export of include/buddha.h.

This is plumbing for tests and bench.py; the product is the shared library and the C command line
(csrc/cudabrot_main.c).  There is no CPU fallback: importing works anywhere (so the symbol table
can be checked without a GPU), but creating a Renderer needs an sm_100 device.
"""
import ctypes as C
import os
import subprocess

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BUDDHA_LIB") or os.path.join(PKG_DIR, "libbuddha.so")  # override: A/B runs
CLI_PATH = os.path.join(PKG_DIR, "bin", "cudabrot")
CSRC_DIR = os.path.join(PKG_DIR, "csrc")

F_NO_SHORTCUT = 1 << 0
F_SIMPLE_KERNEL = 1 << 1
F_EXACT_BINNING = 1 << 2
F_FORCE_TILED = 1 << 3
F_BURNING_SHIP = 1 << 4

ERRORS = {0: "OK", 1: "EINVAL", 2: "ECUDA", 3: "ENOMEM", 4: "ESIZE", 5: "ENCCL", 6: "ENODEV"}

# every symbol include/buddha.h declares (tests check that the library exports exactly these)
EXPORTS = [
    "buddha_abi_version", "buddha_default_params", "buddha_validate_canvas", "buddha_create",
    "buddha_destroy", "buddha_last_error", "buddha_clear_histogram", "buddha_load_histogram",
    "buddha_read_histogram", "buddha_render_samples", "buddha_render_samples_async", "buddha_sync",
    "buddha_render_seconds", "buddha_last_render_ms", "buddha_get_counters",
    "buddha_reset_counters", "buddha_tonemap_u16", "buddha_last_tonemap_ms",
    "buddha_device_histogram", "buddha_stream", "buddha_merge", "buddha_probe_fp64_peak",
    "buddha_probe_red_peak", "buddha_read_channel", "buddha_tonemap_channel_u16",
    "buddha_get_channel_counters", "buddha_device_histogram_cells",
    "buddha_add_histogram_async", "buddha_snapshot", "buddha_read_snapshot",
    "buddha_tonemap_snapshot_u16", "buddha_histogram_digest", "buddha_combine_rgb_u16",
]
MAX_CHANNELS = 4


class Params(C.Structure):
    """buddha_params (include/buddha.h)."""
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32),
                ("min_real", C.c_double), ("max_real", C.c_double),
                ("min_imag", C.c_double), ("max_imag", C.c_double),
                ("max_iterations", C.c_int32), ("min_iterations", C.c_int32),
                ("seed", C.c_uint64), ("flags", C.c_uint32), ("reserved", C.c_uint32),
                ("n_channels", C.c_uint32), ("channel_max", C.c_int32 * 4),
                ("channel_min", C.c_int32 * 4), ("reserved2", C.c_uint32)]


class Counters(C.Structure):
    """buddha_counters (include/buddha.h)."""
    _fields_ = [(n, C.c_uint64) for n in
                ("candidates", "rejected", "hit_max", "too_early", "accepted", "escape_iters",
                 "orbit_points", "increments", "executed_iters", "shortcut_hits",
                 "kernel_launches", "exact_bins")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class BuddhaError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("%s: %s" % (ERRORS.get(code, code), message))
        self.code = code


def build(force=False):
    """Compile libbuddha.so and the cudabrot CLI for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC_DIR, f) for f in
            ("buddha_api.cu", "buddha_kernels.cuh", "cudabrot_main.c", "Makefile")]
    srcs.append(os.path.join(os.path.dirname(PKG_DIR), "include", "buddha.h"))
    newest = max(os.path.getmtime(s) for s in srcs)
    stale = force or not os.path.exists(LIB_PATH) or not os.path.exists(CLI_PATH) \
        or os.path.getmtime(LIB_PATH) < newest or os.path.getmtime(CLI_PATH) < newest
    if stale:
        r = subprocess.run(["make", "-C", CSRC_DIR], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libbuddha.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


_lib = None


def lib():
    """Load libbuddha.so (fails loudly if it has not been built) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libbuddha.so is missing: run `python -c 'import __graft_entry__ as g; "
                          "g.build()'` or `make -C cudabrot_b200/csrc` (there is no fallback path)")
    L = C.CDLL(LIB_PATH)
    ctx = C.c_void_p
    u32p, u16p, u64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint16), C.POINTER(C.c_uint64)
    dblp = C.POINTER(C.c_double)
    L.buddha_abi_version.restype = C.c_uint32
    L.buddha_default_params.argtypes = [C.POINTER(Params)]
    L.buddha_default_params.restype = None
    L.buddha_validate_canvas.argtypes = [C.POINTER(Params), dblp, dblp, C.POINTER(C.c_char_p)]
    L.buddha_create.argtypes = [C.POINTER(ctx), C.POINTER(Params)]
    L.buddha_destroy.argtypes = [ctx]
    L.buddha_destroy.restype = None
    L.buddha_last_error.argtypes = [ctx]
    L.buddha_last_error.restype = C.c_char_p
    L.buddha_clear_histogram.argtypes = [ctx]
    L.buddha_load_histogram.argtypes = [ctx, C.c_void_p, C.c_size_t]
    L.buddha_read_histogram.argtypes = [ctx, C.c_void_p, C.c_size_t]
    L.buddha_render_samples.argtypes = [ctx, C.c_uint64, C.c_uint64]
    L.buddha_render_samples_async.argtypes = [ctx, C.c_uint64, C.c_uint64]
    L.buddha_sync.argtypes = [ctx]
    L.buddha_render_seconds.argtypes = [ctx, C.c_double, C.POINTER(C.c_int), C.c_uint64, u64p, u64p]
    L.buddha_last_render_ms.argtypes = [ctx, C.POINTER(C.c_float)]
    L.buddha_get_counters.argtypes = [ctx, C.POINTER(Counters)]
    L.buddha_reset_counters.argtypes = [ctx]
    L.buddha_tonemap_u16.argtypes = [ctx, C.c_double, C.c_int, C.c_void_p, C.c_size_t, u32p, dblp]
    L.buddha_tonemap_channel_u16.argtypes = [ctx, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                             C.c_size_t, u32p, dblp]
    L.buddha_read_channel.argtypes = [ctx, C.c_int, C.c_void_p, C.c_size_t]
    L.buddha_get_channel_counters.argtypes = [ctx, C.c_int, C.POINTER(Counters)]
    L.buddha_last_tonemap_ms.argtypes = [ctx, C.POINTER(C.c_float)]
    L.buddha_device_histogram.argtypes = [ctx]
    L.buddha_device_histogram.restype = C.c_void_p
    L.buddha_device_histogram_cells.argtypes = [ctx]
    L.buddha_device_histogram_cells.restype = C.c_size_t
    L.buddha_stream.argtypes = [ctx]
    L.buddha_stream.restype = C.c_void_p
    L.buddha_merge.argtypes = [C.POINTER(ctx), C.c_int, C.c_int]
    L.buddha_probe_fp64_peak.argtypes = [ctx, dblp]
    L.buddha_probe_red_peak.argtypes = [ctx, C.c_size_t, dblp]
    L.buddha_add_histogram_async.argtypes = [ctx, C.c_void_p, C.c_size_t]
    L.buddha_snapshot.argtypes = [ctx]
    L.buddha_read_snapshot.argtypes = [ctx, C.c_void_p, C.c_size_t]
    L.buddha_tonemap_snapshot_u16.argtypes = [ctx, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                              C.c_size_t, u32p, dblp]
    L.buddha_histogram_digest.argtypes = [ctx, C.c_int, u64p]
    L.buddha_combine_rgb_u16.argtypes = [ctx, C.POINTER(C.c_int), C.c_double, C.c_int, C.c_double,
                                         C.c_int, C.c_void_p, C.c_size_t, u32p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("buddha_abi_version",):
            fn.restype = C.c_int
    _lib = L
    return L


def default_params():
    p = Params()
    lib().buddha_default_params(C.byref(p))
    return p


def validate_canvas(p):
    """Returns (ok, delta_real, delta_imag, message) -- RecomputePixelDeltas, cudabrot.cu:505-527."""
    dr, di, why = C.c_double(), C.c_double(), C.c_char_p()
    rc = lib().buddha_validate_canvas(C.byref(p), C.byref(dr), C.byref(di), C.byref(why))
    return rc == 0, dr.value, di.value, (why.value.decode() if why.value else None)
