"""ctypes binding of oracle/liboracle.so -- the CPU checker (test infrastructure only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package (cudabrot_b200) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")
REF_PROBE = os.path.join(REF_DIR, "ref_probe")
REF_PROBE_SHIP = os.path.join(REF_DIR, "ref_probe_ship")  # built with -DRENDER_BURNING_SHIP
REF_BINARY = os.path.join(REF_DIR, "cudabrot_ref")


class Dims(C.Structure):
    """FractalDimensions, cudabrot.cu:46-58."""
    _fields_ = [("w", C.c_int32), ("h", C.c_int32),
                ("min_real", C.c_double), ("min_imag", C.c_double),
                ("max_real", C.c_double), ("max_imag", C.c_double),
                ("delta_real", C.c_double), ("delta_imag", C.c_double)]


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in
                ("candidates", "rejected", "hit_max", "too_early", "accepted",
                 "escape_iters", "orbit_points", "increments")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(ORACLE_DIR, "buddha_oracle.c")):
        subprocess.run(["make", "-C", ORACLE_DIR, "liboracle.so"], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        u32p, i32p, u16p = (C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_uint16))
        L.oracle_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.oracle_sample.argtypes = [C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.oracle_set_deltas.argtypes = [C.POINTER(Dims)]
        L.oracle_set_deltas.restype = C.c_int
        L.oracle_rejected.argtypes = [C.c_double, C.c_double]
        L.oracle_rejected.restype = C.c_int
        L.oracle_escape_iterations.argtypes = [C.c_double, C.c_double, C.c_int]
        L.oracle_escape_iterations.restype = C.c_int
        L.oracle_escape_iterations_ship.argtypes = [C.c_double, C.c_double, C.c_int]
        L.oracle_escape_iterations_ship.restype = C.c_int
        L.oracle_cycle_detect_iterations.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
        L.oracle_cycle_detect_iterations.restype = C.c_int
        L.oracle_render.argtypes = [C.POINTER(Dims), C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                    C.c_uint64, u32p, C.POINTER(Counters), C.c_int]
        L.oracle_render.restype = C.c_int
        L.oracle_render_ex.argtypes = L.oracle_render.argtypes + [C.c_int]
        L.oracle_render_ex.restype = C.c_int
        L.oracle_classify.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, i32p]
        L.oracle_classify_ship.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, i32p]
        L.oracle_check_scaled_ship.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.oracle_check_scaled_ship.restype = C.c_uint64
        L.oracle_tonemap.argtypes = [u32p, C.c_size_t, C.c_double, C.c_int, u16p, u32p,
                                     C.POINTER(C.c_double)]
        L.oracle_write_pgm.argtypes = [C.c_char_p, u16p, C.c_int, C.c_int]
        L.oracle_write_pgm.restype = C.c_int
        L.oracle_fnv1a64.argtypes = [u32p, C.c_size_t]
        L.oracle_fnv1a64.restype = C.c_uint64
        L.oracle_max_threads.restype = C.c_int
        L.oracle_escape_iterations_scaled.argtypes = [C.c_double, C.c_double, C.c_int]
        L.oracle_escape_iterations_scaled.restype = C.c_int
        L.oracle_check_scaled.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
        L.oracle_check_scaled.restype = C.c_uint64
        L.oracle_check_fast_bin.argtypes = [C.POINTER(Dims), C.POINTER(C.c_double), C.c_uint64,
                                            C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.oracle_check_fast_bin.restype = C.c_uint64
        L.oracle_bin_fast.argtypes = [C.POINTER(Dims), C.c_double, C.c_double,
                                      C.POINTER(C.c_int64), C.POINTER(C.c_int)]
        L.oracle_bin_fast.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    o = np.zeros(4, dtype=np.uint32)
    lib().oracle_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(o, C.c_uint32))
    return o


def sample(seed, s):
    re, im = C.c_double(), C.c_double()
    lib().oracle_sample(seed, s, C.byref(re), C.byref(im))
    return re.value, im.value


def samples(seed, first, count):
    out = np.empty((count, 2), dtype=np.float64)
    for k in range(count):
        out[k] = sample(seed, first + k)
    return out


def make_dims(w, h, min_real=-2.0, max_real=2.0, min_imag=-2.0, max_imag=2.0):
    d = Dims(w, h, min_real, min_imag, max_real, max_imag, 0.0, 0.0)
    if not lib().oracle_set_deltas(C.byref(d)):
        raise ValueError("invalid canvas")
    return d


def render(w, h, max_iter, min_iter, seed, first, count, canvas=(-2.0, 2.0, -2.0, 2.0),
           hist=None, threads=0, burning_ship=False):
    """Returns (hist uint32[h,w], counters dict, threads used)."""
    d = make_dims(w, h, *canvas)
    if hist is None:
        hist = np.zeros((h, w), dtype=np.uint32)
    cnt = Counters()
    nt = lib().oracle_render_ex(C.byref(d), max_iter, min_iter, seed, first, count,
                                _p(hist, C.c_uint32), C.byref(cnt), threads, int(burning_ship))
    return hist, cnt.as_dict(), nt


def classify(seed, first, count, max_iter, burning_ship=False):
    out = np.empty(count, dtype=np.int32)
    fn = lib().oracle_classify_ship if burning_ship else lib().oracle_classify
    fn(seed, first, count, max_iter, _p(out, C.c_int32))
    return out


def tonemap(hist, gamma, big_endian=False):
    """Returns (uint16 image (same shape), max, scale)."""
    h = np.ascontiguousarray(hist, dtype=np.uint32)
    out = np.empty(h.shape, dtype=np.uint16)
    mx, sc = C.c_uint32(), C.c_double()
    lib().oracle_tonemap(_p(h, C.c_uint32), h.size, gamma, int(big_endian), _p(out, C.c_uint16),
                         C.byref(mx), C.byref(sc))
    return out, mx.value, sc.value


def write_pgm(path, image):
    img = np.ascontiguousarray(image, dtype=np.uint16)
    rc = lib().oracle_write_pgm(path.encode(), _p(img, C.c_uint16), img.shape[1], img.shape[0])
    if rc:
        raise OSError("oracle_write_pgm failed: %d" % rc)


def fnv1a64(hist):
    h = np.ascontiguousarray(hist, dtype=np.uint32)
    return int(lib().oracle_fnv1a64(_p(h, C.c_uint32), h.size))


def blocked_fnv(hist):
    """The digest buddha_histogram_digest forms on the GPU (include/buddha.h), restated in numpy:
    blocks of 4096 cells; lane l = 0..31 of a block folds cells l, l+32, ... with FNV-1a-64 over
    whole cells; the 32 lane values fold into the block digest, the block digests into the result."""
    basis, prime = np.uint64(0xcbf29ce484222325), np.uint64(0x100000001b3)
    cells = np.ascontiguousarray(hist, dtype=np.uint32).reshape(-1)
    n_blocks = (cells.size + 4095) // 4096
    padded = np.zeros(n_blocks * 4096, dtype=np.uint64)
    padded[:cells.size] = cells
    v = padded.reshape(n_blocks, 128, 32)
    with np.errstate(over="ignore"):
        h = np.full((n_blocks, 32), basis, dtype=np.uint64)
        for k in range(128):
            h = (h ^ v[:, k, :]) * prime
        d = np.full(n_blocks, basis, dtype=np.uint64)
        for l in range(32):
            d = (d ^ h[:, l]) * prime
        out = basis
        for b in range(n_blocks):
            out = (out ^ d[b]) * prime
    return int(out)


def check_period3(seed, first, count, max_iter, limit=0.998):
    """(flagged samples that escaped -- must be 0, flagged, all never-escaping samples): the evidence
    behind the kernel's conservative period-3 test (buddha_kernels.cuh: in_period3_component)."""
    L = lib()
    L.oracle_check_period3.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_float,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.oracle_check_period3.restype = C.c_uint64
    fl, ins = C.c_uint64(), C.c_uint64()
    bad = L.oracle_check_period3(seed, first, count, max_iter, limit, C.byref(fl), C.byref(ins))
    return int(bad), fl.value, ins.value


def check_period4(seed, first, count, max_iter, mu_max=0.999 / 16):
    """Like check_period3 for the period-4 test (buddha_kernels.cuh: in_period4_component)."""
    L = lib()
    L.oracle_check_period4.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_float,
                                       C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    L.oracle_check_period4.restype = C.c_uint64
    fl, ins = C.c_uint64(), C.c_uint64()
    bad = L.oracle_check_period4(seed, first, count, max_iter, mu_max, C.byref(fl), C.byref(ins))
    return int(bad), fl.value, ins.value


def check_certificate(seed, first, count, max_iter, lam2_max=0.998, pmax=(32, 64), tol=1e-2, passes=5,
                      res_max=1e-12, first_age=64, age_factor=4):
    """(certified samples that escaped -- must be 0, stats): the evidence behind the kernel's
    attracting-cycle certificate (buddha_kernels.cuh: cert_phase), tried where the kernel tries it.
    stats: not_rejected, inset (never-escaping, not flagged by the period-3/4 tests), certified,
    iterations of the inset samples with the bit-exact search only / with the certificate,
    attempts."""
    L = lib()
    L.oracle_check_certificate.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.c_int,
                                           C.c_int, C.c_int, C.c_float, C.c_int, C.c_double,
                                           C.c_double, C.POINTER(C.c_uint64)]
    L.oracle_check_certificate.restype = C.c_uint64
    st = (C.c_uint64 * 8)()
    bad = L.oracle_check_certificate(seed, first, count, max_iter, first_age, age_factor, pmax[0],
                                     pmax[1], tol, passes, res_max, lam2_max, st)
    keys = ("not_rejected", "inset", "certified", "iters_exact", "iters_cert", "attempts",
            "searched", "periods")
    return int(bad), {k: int(v) for k, v in zip(keys, st)}


def check_prefilter(seed, first, count, ship=False, m_rej=0.02, m_esc=0.05):
    """(decided samples that disagree with the reference's arithmetic -- must be 0, [undecided,
    rejected, escapes at step 1, at step 2]): the evidence behind the sampler's FP32
    pre-classification (buddha_kernels.cuh: gen_phase)."""
    L = lib()
    L.oracle_check_prefilter.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_float,
                                         C.c_float, C.POINTER(C.c_uint64)]
    L.oracle_check_prefilter.restype = C.c_uint64
    cnt = (C.c_uint64 * 4)()
    bad = L.oracle_check_prefilter(seed, first, count, int(ship), m_rej, m_esc, cnt)
    return int(bad), [int(x) for x in cnt]


def max_threads():
    return int(lib().oracle_max_threads())


def check_scaled(seed, first, count, max_iter, burning_ship=False):
    """Mismatches between the reference-form and the product's scaled recurrence / rejection."""
    fn = lib().oracle_check_scaled_ship if burning_ship else lib().oracle_check_scaled
    return int(fn(seed, first, count, max_iter))


def check_fast_bin(dims, points):
    """(mismatches, points that took the division path, points in canvas); mismatches is None
    when the canvas does not admit the division-free path."""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    idx, took = C.c_int64(), C.c_int()
    if lib().oracle_bin_fast(C.byref(dims), 0.0, 0.0, C.byref(idx), C.byref(took)) < 0:
        return None, 0, 0
    ex, inc = C.c_uint64(), C.c_uint64()
    bad = lib().oracle_check_fast_bin(C.byref(dims), _p(pts, C.c_double), len(pts), C.byref(ex),
                                      C.byref(inc))
    return int(bad), int(ex.value), int(inc.value)


def run_ref_probe(*args, burning_ship=False):
    """Run oracle/_ref/ref_probe (the reference's own code, prebuilt); returns CompletedProcess."""
    exe = REF_PROBE_SHIP if burning_ship else REF_PROBE
    return subprocess.run([exe] + [str(a) for a in args], capture_output=True, text=True)
