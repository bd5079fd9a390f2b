"""CPU tests of the drop-in boundary: the C-ABI library loads, exports exactly what
include/buddha.h declares, validates like the reference, and refuses to run without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "buddha.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(buddha_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(buddha):
    declared = _declared_symbols()
    assert declared == sorted(buddha.capi.EXPORTS)
    nm = subprocess.run(["nm", "-D", "--defined-only", buddha.capi.LIB_PATH], capture_output=True,
                        text=True, check=True).stdout
    exported = sorted(set(re.findall(r" T (buddha_[a-z0-9_]+)", nm)))
    assert exported == declared
    L = buddha.capi.lib()
    for name in declared:
        assert getattr(L, name) is not None
    assert L.buddha_abi_version() == 3


def test_no_torch_or_oracle_dependency(buddha):
    ldd = subprocess.run(["ldd", buddha.capi.LIB_PATH], capture_output=True, text=True).stdout
    # library names only: the load addresses ldd prints may contain "c10" by chance
    names = " ".join(line.split("=>")[0].split("(")[0].strip() for line in ldd.splitlines())
    assert "torch" not in names and "oracle" not in names and "c10" not in names
    for src in ("buddha_api.cu", "buddha_kernels.cuh", "cudabrot_main.c"):
        text = open(os.path.join(ROOT, "cudabrot_b200", "csrc", src)).read()
        assert "oracle" not in text.lower(), src


def test_default_params_are_the_reference_defaults(buddha):
    p = buddha.capi.default_params()   # cudabrot.cu:763-772, :530-543
    assert (p.width, p.height, p.max_iterations, p.min_iterations) == (1000, 1000, 100, 20)
    assert (p.min_real, p.max_real, p.min_imag, p.max_imag) == (-2.0, 2.0, -2.0, 2.0)
    assert p.seed == 1337 and p.device == 0 and p.struct_size == C.sizeof(buddha.capi.Params)


def test_validate_canvas_rules_and_deltas(buddha, oracle):
    p = buddha.capi.default_params()
    ok, dr, di, why = buddha.capi.validate_canvas(p)
    assert ok and dr == 4.0 / 1000 and di == 4.0 / 1000 and why is None
    # deltas are computed exactly like RecomputePixelDeltas (cudabrot.cu:524-525)
    p.width, p.height = 777, 333
    p.min_real, p.max_real, p.min_imag, p.max_imag = -1.7, 0.3, -0.123, 0.777
    ok, dr, di, _ = buddha.capi.validate_canvas(p)
    d = oracle.make_dims(777, 333, -1.7, 0.3, -0.123, 0.777)
    assert ok and dr == d.delta_real and di == d.delta_imag
    cases = [("width", 0, "Output width must be positive."),
             ("height", -3, "Output height must be positive."),
             ("max_real", -1.7, "Maximum real value must be greater than minimum real value."),
             ("max_imag", -5.0,
              "Minimum imaginary value must be greater than maximum imaginary value.")]
    for field, value, msg in cases:
        q = buddha.capi.default_params()
        q.min_real = -1.7
        setattr(q, field, value)
        ok, _, _, why = buddha.capi.validate_canvas(q)
        assert not ok and why == msg


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(buddha):
    with pytest.raises(buddha.BuddhaError) as e:
        buddha.Renderer(64, 64)
    assert "no CPU fallback" in str(e.value)


def test_create_rejects_bad_params(buddha):
    L = buddha.capi.lib()
    ctx = C.c_void_p()
    p = buddha.capi.default_params()
    p.struct_size = 8
    assert L.buddha_create(C.byref(ctx), C.byref(p)) == 1
    p = buddha.capi.default_params()
    p.width = 0
    assert L.buddha_create(C.byref(ctx), C.byref(p)) == 1
    assert b"width" in L.buddha_last_error(None)
    p = buddha.capi.default_params()
    p.width, p.height = 50000, 50000          # > 2^31-1 cells (32-bit index, cudabrot.cu:312)
    assert L.buddha_create(C.byref(ctx), C.byref(p)) == 1
