"""Renderer: a thin object over one libbuddha context (one GPU).

Mirrors the reference program's flow (cudabrot.cu:762-791): SetupCUDA -> LoadInProgressBuffer ->
RenderImage -> SaveInProgressBuffer -> SaveImage, with each stage a method that calls straight
through the C ABI.  Host buffers are numpy arrays; nothing here computes anything.
"""
import ctypes as C

import numpy as np

from . import capi


class Renderer:
    def __init__(self, width=1000, height=1000, max_iterations=100, min_iterations=20,
                 canvas=(-2.0, 2.0, -2.0, 2.0), seed=1337, device=0, flags=0, channels=None):
        """canvas = (min_real, max_real, min_imag, max_imag).  Defaults = the reference's defaults
        (cudabrot.cu:763-772, :530-543).  channels = [(max_iterations, min_iterations), ...] (2..4
        entries) makes a fused multi-channel context: every candidate is rendered once into each
        channel whose window accepts it (what generate_hires_color_image.sh:27-59 does with one
        run of the reference per channel)."""
        self._lib = capi.lib()
        p = capi.default_params()
        p.device = device
        p.width, p.height = width, height
        p.min_real, p.max_real, p.min_imag, p.max_imag = canvas
        p.max_iterations, p.min_iterations = max_iterations, min_iterations
        p.seed, p.flags = seed, flags
        self.channels = list(channels) if channels else None
        if self.channels is not None and len(self.channels) == 1:
            # (one pair would silently fall back to max_iterations / min_iterations)
            raise ValueError("channels needs 2..%d (max, min) pairs; use max_iterations / "
                             "min_iterations for a single channel" % capi.MAX_CHANNELS)
        if self.channels:
            p.n_channels = len(self.channels)
            for k, (m, c) in enumerate(self.channels[:capi.MAX_CHANNELS]):
                p.channel_max[k], p.channel_min[k] = m, c
        self.n_channels = len(self.channels) if self.channels else 1
        self.params = p
        self.width, self.height = width, height
        self.cells = width * height * self.n_channels
        self._ctx = C.c_void_p()
        rc = self._lib.buddha_create(C.byref(self._ctx), C.byref(p))
        if rc:
            msg = self._lib.buddha_last_error(None).decode()
            self._ctx = None
            raise capi.BuddhaError(rc, msg)

    # -- lifetime ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.buddha_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        self.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc):
        if rc:
            raise capi.BuddhaError(rc, self._lib.buddha_last_error(self._ctx).decode())

    # -- histogram (the -s buffer surface, cudabrot.cu:215-280) -------------------------------
    def clear(self):
        self._check(self._lib.buddha_clear_histogram(self._ctx))

    def load_histogram(self, hist):
        h = np.ascontiguousarray(hist, dtype=np.uint32)
        self._check(self._lib.buddha_load_histogram(self._ctx, h.ctypes.data, h.size))

    def read_histogram(self, out=None):
        """uint32[h, w]; fused contexts: uint32[n_channels, h, w]."""
        if out is None:
            shape = (self.height, self.width)
            out = np.empty(((self.n_channels,) + shape) if self.channels else shape, dtype=np.uint32)
        self._check(self._lib.buddha_read_histogram(self._ctx, out.ctypes.data, out.size))
        return out

    def read_channel(self, channel, out=None):
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint32)
        self._check(self._lib.buddha_read_channel(self._ctx, channel, out.ctypes.data, out.size))
        return out

    # -- render (cudabrot.cu:379-414, :483-492) -----------------------------------------------
    def render_samples(self, first, count):
        self._check(self._lib.buddha_render_samples(self._ctx, first, count))

    def render_samples_async(self, first, count):
        self._check(self._lib.buddha_render_samples_async(self._ctx, first, count))

    def sync(self):
        self._check(self._lib.buddha_sync(self._ctx))

    def render_seconds(self, seconds, first=0, stop=None):
        """stop: an optional ctypes.c_int another thread may set to end the run (the reference's
        quit_signal_received, cudabrot.cu:483); required when seconds < 0."""
        done, passes = C.c_uint64(), C.c_uint64()
        self._check(self._lib.buddha_render_seconds(self._ctx, seconds,
                                                    C.byref(stop) if stop is not None else None,
                                                    first, C.byref(done), C.byref(passes)))
        return done.value, passes.value

    def last_render_ms(self):
        ms = C.c_float()
        self._check(self._lib.buddha_last_render_ms(self._ctx, C.byref(ms)))
        return ms.value

    def counters(self):
        c = capi.Counters()
        self._check(self._lib.buddha_get_counters(self._ctx, C.byref(c)))
        return c.as_dict()

    def channel_counters(self, channel):
        c = capi.Counters()
        self._check(self._lib.buddha_get_channel_counters(self._ctx, channel, C.byref(c)))
        return c.as_dict()

    def reset_counters(self):
        self._check(self._lib.buddha_reset_counters(self._ctx))

    # -- tone-map (cudabrot.cu:416-468, :566-570) ----------------------------------------------
    def tonemap(self, gamma=1.0, big_endian=False, out=None, want_image=True, channel=0):
        """Returns (uint16 image or None, max, scale)."""
        mx, sc = C.c_uint32(), C.c_double()
        ptr, n = None, 0
        if want_image:
            if out is None:
                out = np.empty((self.height, self.width), dtype=np.uint16)
            ptr, n = out.ctypes.data, out.size
        self._check(self._lib.buddha_tonemap_channel_u16(self._ctx, channel, gamma,
                                                         int(big_endian), ptr, n, C.byref(mx),
                                                         C.byref(sc)))
        return (out if want_image else None), mx.value, sc.value

    def last_tonemap_ms(self):
        ms = C.c_float()
        self._check(self._lib.buddha_last_tonemap_ms(self._ctx, C.byref(ms)))
        return ms.value

    # -- overlapped host transfers (no reference equivalent: its copies block) ----------------
    def add_histogram_async(self, hist):
        """Adds host counts to the histogram; the copy overlaps whatever is rendering.  `hist`
        must stay alive (and should be page-locked) until sync()."""
        if hist.dtype != np.uint32 or not hist.flags["C_CONTIGUOUS"]:
            raise TypeError("add_histogram_async needs a C-contiguous uint32 array (no copy is made)")
        self._check(self._lib.buddha_add_histogram_async(self._ctx, hist.ctypes.data, hist.size))

    def snapshot(self):
        self._check(self._lib.buddha_snapshot(self._ctx))

    def read_snapshot(self, out=None):
        if out is None:
            shape = (self.height, self.width)
            out = np.empty(((self.n_channels,) + shape) if self.channels else shape, dtype=np.uint32)
        self._check(self._lib.buddha_read_snapshot(self._ctx, out.ctypes.data, out.size))
        return out

    def tonemap_snapshot(self, gamma=1.0, big_endian=False, out=None, channel=0):
        """Returns (uint16 image, max, scale) of the snapshot."""
        mx, sc = C.c_uint32(), C.c_double()
        if out is None:
            out = np.empty((self.height, self.width), dtype=np.uint16)
        self._check(self._lib.buddha_tonemap_snapshot_u16(self._ctx, channel, gamma,
                                                          int(big_endian), out.ctypes.data,
                                                          out.size, C.byref(mx), C.byref(sc)))
        return out, mx.value, sc.value

    def combine_rgb(self, channels=(0, 1, 2), gamma=1.0, mode="rgb", hue_adjust=0.0,
                    big_endian=False, out=None):
        """uint16[h, w, 3] colour image from three channels, combined on the GPU; mode "rgb" or
        "hsl" (hue, saturation, lightness = the three channels).  Returns (image, [max per channel])."""
        if out is None:
            out = np.empty((self.height, self.width, 3), dtype=np.uint16)
        ch = (C.c_int * 3)(*channels)
        mx = (C.c_uint32 * 3)()
        self._check(self._lib.buddha_combine_rgb_u16(self._ctx, ch, gamma,
                                                     {"rgb": 0, "hsl": 1}[mode], hue_adjust,
                                                     int(big_endian), out.ctypes.data,
                                                     self.width * self.height, mx))
        return out, [int(v) for v in mx]

    def digest(self, channel=0):
        """64-bit blocked FNV-1a digest of one channel, formed on the GPU (include/buddha.h)."""
        d = C.c_uint64()
        self._check(self._lib.buddha_histogram_digest(self._ctx, channel, C.byref(d)))
        return d.value

    # -- multi-GPU plumbing --------------------------------------------------------------------
    @property
    def device_histogram_ptr(self):
        return self._lib.buddha_device_histogram(self._ctx)

    @property
    def stream_ptr(self):
        return self._lib.buddha_stream(self._ctx)

    def histogram_as_tensor(self):
        """The device histogram as a torch int32 tensor aliasing libbuddha's memory (no copy), so
        torch.distributed can reduce it in place.  torch is plumbing here, nothing more."""
        import torch

        class _Alias:
            pass

        a = _Alias()
        a.__cuda_array_interface__ = {
            "shape": (int(self._lib.buddha_device_histogram_cells(self._ctx)),), "typestr": "<i4",
            "data": (self.device_histogram_ptr, False),
            "version": 2, "strides": None,
        }
        return torch.as_tensor(a, device="cuda:%d" % self.params.device)

    # -- roofline probes -----------------------------------------------------------------------
    def probe_fp64_peak(self):
        v = C.c_double()
        self._check(self._lib.buddha_probe_fp64_peak(self._ctx, C.byref(v)))
        return v.value

    def probe_red_peak(self, footprint_bytes):
        v = C.c_double()
        self._check(self._lib.buddha_probe_red_peak(self._ctx, footprint_bytes, C.byref(v)))
        return v.value


def merge_in_process(renderers, root=0):
    """buddha_merge: sum the histograms of several in-process contexts into renderers[root]."""
    L = capi.lib()
    arr = (C.c_void_p * len(renderers))(*[r._ctx for r in renderers])
    rc = L.buddha_merge(arr, len(renderers), root)
    if rc:
        raise capi.BuddhaError(rc, L.buddha_last_error(renderers[root]._ctx).decode())


def write_pgm(path, image_be, width, height):
    """SaveImage (cudabrot.cu:548-577) for an image that is already big-endian."""
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n%d\n" % (width, height, 0xffff))
        f.write(np.ascontiguousarray(image_be, dtype=np.uint16).tobytes())
