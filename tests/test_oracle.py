"""CPU tests of the oracle: pins the CPU restatement against every known answer available without
a GPU.  The reference ships no tests or golden vectors of its own."""
import glob
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# Random123 known-answer vectors for Philox4x32-10 (kat_vectors of the Random123 distribution;
# cuRAND's curand_Philox4x32_10, curand_philox4x32_x.h:159-192, implements the same function).
@pytest.mark.parametrize("ctr,key,expect", [
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
])
def test_philox_known_answers(oracle, ctr, key, expect):
    assert [int(v) for v in oracle.philox(ctr, key)] == expect


def test_sample_mapping(oracle):
    # seed 1337, samples 0 and 1 (words and resulting c)
    assert [int(v) for v in oracle.philox([0, 0, 0, 0], [1337, 0])] == \
        [0xdab1c1f7, 0xf009740a, 0xdb39b086, 0xdc0d2cde]
    re, im = oracle.sample(1337, 0)
    assert re == float.fromhex("0x1.c025cb7e383fp+0") and im == float.fromhex("0x1.7034a81f3610cp+0")
    re, im = oracle.sample(1337, 1)
    assert re == float.fromhex("0x1.cecc997e34e34p+0") and im == float.fromhex("-0x1.c5c3d7e7ee256p-1")
    # samples cover (-2, 2]: cudabrot.cu:392-393 with u in (0, 1]
    pts = oracle.samples(7, 0, 2000)
    assert pts.min() > -2.0 and pts.max() <= 2.0
    # 64-bit sample indices use both counter words
    assert oracle.sample(1337, 1 << 32) != oracle.sample(1337, 0)


# (w, h, canvas, m, c, N) -> counters measured by the survey's independent probe
SURVEY_KATS = [
    ((64, 64, (-2.0, 2.0, -2.0, 2.0), 100, 20, 1 << 16),
     dict(accepted=680, orbit_points=27332, increments=26784), 86, 2105),
    ((200, 100, (0.0, 1.0, 0.0, 0.5), 1000, 20, 1 << 20),
     dict(accepted=13726, orbit_points=1060270, increments=98503), 33, 14867),
    ((1000, 1000, (-2.0, 2.0, -2.0, 2.0), 100, 20, 1 << 24),
     dict(accepted=180775, orbit_points=7231409, increments=7085954), 372, 525289),
]


@pytest.mark.parametrize("cfg,counters,hmax,nonzero", SURVEY_KATS)
def test_render_matches_survey_probe(oracle, cfg, counters, hmax, nonzero):
    w, h, canvas, m, c, n = cfg
    hist, cnt, _ = oracle.render(w, h, m, c, 1337, 0, n, canvas=canvas)
    for k, v in counters.items():
        assert cnt[k] == v, k
    assert int(hist.max()) == hmax
    assert int(np.count_nonzero(hist)) == nonzero
    assert int(hist.sum()) == cnt["increments"]
    assert cnt["candidates"] == n
    assert cnt["rejected"] + cnt["hit_max"] + cnt["too_early"] + cnt["accepted"] == n


def test_render_is_thread_and_split_invariant(oracle):
    a, ca, _ = oracle.render(128, 96, 200, 10, 99, 5, 50000, threads=1)
    b, cb, _ = oracle.render(128, 96, 200, 10, 99, 5, 50000, threads=4)
    assert np.array_equal(a, b) and ca == cb
    c1, _, _ = oracle.render(128, 96, 200, 10, 99, 5, 20000)
    c2, _, _ = oracle.render(128, 96, 200, 10, 99, 20005, 30000, hist=c1)
    assert np.array_equal(a, c2)


def test_quirks(oracle):
    # -m <= 0: nothing is ever accepted
    hist, cnt, _ = oracle.render(32, 32, 0, 0, 1337, 0, 4096)
    assert hist.sum() == 0 and cnt["accepted"] == 0 and cnt["escape_iters"] == 0
    assert cnt["hit_max"] + cnt["rejected"] == 4096
    # z0 = c is not recorded, the escaped point is (quirk 2): orbit points = i + 1
    it = oracle.classify(1337, 0, 4096, 50)
    hist, cnt, _ = oracle.render(32, 32, 50, 0, 1337, 0, 4096)
    assert cnt["orbit_points"] == int(((it[(it >= 0) & (it < 50)]) + 1).sum())
    # invalid canvases (cudabrot.cu:505-523)
    for bad in [(0, 10, -2, 2, -2, 2), (10, -1, -2, 2, -2, 2), (10, 10, 2, 2, -2, 2),
                (10, 10, -2, 2, 1, 0)]:
        with pytest.raises(ValueError):
            oracle.make_dims(*bad)


def test_burning_ship_differs_and_skips_rejection(oracle):
    """RENDER_BURNING_SHIP (cudabrot.cu:15-17, :397-399): no cardioid/bulb rejection, a different
    set; the upper-right quadrant start (re, im >= 0 throughout) is NOT enough to make the two
    variants agree because the orbit leaves the quadrant."""
    h0, c0, _ = oracle.render(128, 128, 200, 10, 1337, 0, 1 << 16)
    h1, c1, _ = oracle.render(128, 128, 200, 10, 1337, 0, 1 << 16, burning_ship=True)
    assert c1["rejected"] == 0 and c0["rejected"] > 0
    assert c1["candidates"] == c0["candidates"] == 1 << 16
    assert not np.array_equal(h0, h1)
    assert c1["accepted"] > 0 and int(h1.sum()) == c1["increments"]
    # single points: c = -1.45 - 0.1i is inside the ship and escapes after 6 Mandelbrot steps
    L = oracle.lib()
    assert L.oracle_escape_iterations_ship(-1.45, -0.1, 5000) == 5000
    assert L.oracle_escape_iterations(-1.45, -0.1, 5000) == 6


def test_scaled_recurrence_is_bit_identical(oracle):
    """The product's 4-instruction scaled step and scaled cardioid/bulb test give the same escape
    index / verdict as the reference dataflow on every sample."""
    assert oracle.check_scaled(1337, 0, 1 << 22, 500) == 0
    assert oracle.check_scaled(42, 1 << 40, 1 << 20, 5000) == 0
    # the burning-ship variant (|re|, |im| before each step) under the same scaling
    assert oracle.check_scaled(1337, 0, 1 << 21, 500, burning_ship=True) == 0
    assert oracle.check_scaled(7, 1 << 35, 1 << 19, 5000, burning_ship=True) == 0


CANVASES = [(1000, 1000, (-2.0, 2.0, -2.0, 2.0)), (20000, 20000, (-2.0, 2.0, -2.0, 2.0)),
            (8000, 4000, (0.0, 1.0, 0.0, 0.5)), (777, 333, (-1.7, 0.3, -0.123, 0.777)),
            (3, 5, (-2.0, 2.0, -2.0, 2.0)), (100000, 10, (-2.0, 1.0, -0.001, 0.001))]


@pytest.mark.parametrize("w,h,canvas", CANVASES)
def test_division_free_binning_is_bit_identical(oracle, w, h, canvas):
    """orbit_bin's one-rounding test (T = rn(quotient + 2^-11) on a 2^-12 grid, T - 2^-10 < Q < T)
    either reproduces trunc(rn((v-min)/delta)) or defers to the IEEE division -- including on exact
    pixel boundaries, their floating-point neighbours, and the edges of the deferred sliver."""
    d = oracle.make_dims(w, h, canvas[0], canvas[1], canvas[2], canvas[3])
    rng = np.random.default_rng(w * 31 + h)
    n = 400_000
    pts = rng.uniform(-3, 3, size=(n, 2))
    q = n // 4
    bx = canvas[0] + rng.integers(-2, w + 3, size=q) * d.delta_real
    by = canvas[2] + rng.integers(-2, h + 3, size=q) * d.delta_imag
    pts[:q, 0], pts[:q, 1] = bx, np.nextafter(by, -np.inf)
    pts[q:2 * q, 0], pts[q:2 * q, 1] = np.nextafter(bx, np.inf), by
    pts[2 * q:3 * q, 0] = np.nextafter(bx, -np.inf)
    pts[3 * q:, 0] = rng.uniform(canvas[0], canvas[1], size=n - 3 * q)
    pts[3 * q:, 1] = rng.uniform(canvas[2], canvas[3], size=n - 3 * q)
    bad, exact, inside = oracle.check_fast_bin(d, pts)
    assert bad == 0
    assert inside > 0 and exact > 0
    # the edges of the sliver: 2^-12 ... 2^-9 of a pixel on either side of a boundary, +- 1 ulp
    edge = []
    for k in (-9, -10, -11, -12):
        for sign in (1.0, -1.0):
            ex = bx + sign * 2.0 ** k * d.delta_real
            ey = by + sign * 2.0 ** k * d.delta_imag
            for fx in (ex, np.nextafter(ex, np.inf), np.nextafter(ex, -np.inf)):
                edge.append(np.stack([fx, ey], axis=1))
                edge.append(np.stack([fx, rng.uniform(canvas[2], canvas[3], size=q)], axis=1))
    bad3, _, inside3 = oracle.check_fast_bin(d, np.concatenate(edge))
    assert bad3 == 0 and inside3 > 0
    # ordinary interior points almost never need the division
    bad2, exact2, inside2 = oracle.check_fast_bin(d, pts[3 * q:])
    assert bad2 == 0 and exact2 <= 0.01 * inside2 + 50


def test_binning_fast_path_declines_extreme_zoom(oracle):
    d = oracle.make_dims(1000, 1000, -0.743644786, -0.7436447859, 0.1318252536, 0.1318252537)
    assert oracle.check_fast_bin(d, np.zeros((4, 2)))[0] is None


def test_cycle_detection_implies_never_escapes(oracle):
    """A bit-for-bit repeat of z proves the orbit periodic: such samples must be hit_max."""
    it = oracle.classify(1337, 0, 1 << 16, 3000)
    L = oracle.lib()
    seen = 0
    for s in np.nonzero(it >= 0)[0][:20000]:
        re, im = oracle.sample(1337, int(s))
        n = L.oracle_cycle_detect_iterations(re, im, 3000, 8)
        if n > 0:
            seen += 1
            assert it[s] == 3000
    assert seen > 100


def _golden_cases():
    for pgm in sorted(glob.glob(os.path.join(GOLDEN, "tonemap_*.pgm"))):
        name, g = os.path.basename(pgm)[len("tonemap_"):-len(".pgm")].rsplit("_g", 1)
        yield pytest.param(name, float(g), pgm, id="%s-g%s" % (name, g))


@pytest.mark.parametrize("name,gamma,pgm", list(_golden_cases()))
def test_tonemap_and_pgm_match_reference_host_code(oracle, tmp_path, name, gamma, pgm):
    """Golden files were written by the reference's own SetGrayscalePixels + SaveImage
    (tests/golden/make_golden.py via oracle/_ref/ref_probe)."""
    raw = np.fromfile(os.path.join(GOLDEN, "hist_%s.raw" % name), dtype="<u4")
    side = {"kat64": (64, 64), "big": (32, 32), "zero": (16, 16)}[name]
    hist = raw.reshape(side)
    img, mx, scale = oracle.tonemap(hist, gamma)
    out = str(tmp_path / "o.pgm")
    oracle.write_pgm(out, img)
    assert open(out, "rb").read() == open(pgm, "rb").read()
    line = open(pgm + ".stdout").read().strip()
    assert line == "Max value: %d, scale: %f" % (mx, scale)
    # big-endian mode == byte-swapped host-endian mode (cudabrot.cu:566-570)
    be, _, _ = oracle.tonemap(hist, gamma, big_endian=True)
    assert np.array_equal(be, img.byteswap())


def test_tonemap_max_pixel_quirk(oracle):
    # with max = 372 the brightest pixel maps to 65534, not 65535, at gamma 1
    h = np.array([[0, 1, 372]], dtype=np.uint32)
    img, mx, _ = oracle.tonemap(h, 1.0)
    assert mx == 372 and int(img[0, 2]) == 65534
    img, _, _ = oracle.tonemap(h, 2.2)
    assert int(img[0, 2]) == 65535


def test_period3_component_test_is_conservative(oracle):
    """The kernel skips the escape loop for samples whose period-3 multiplier (closed form,
    Giarrusso & Fisher) has |lambda|^2 < 0.998.  Every such sample must run the reference's loop
    (cudabrot.cu:319-340) to max_iterations; a limit beyond the component boundary (|lambda| = 1)
    must be caught by this very check."""
    for seed, first in ((1337, 0), (99, 1 << 40)):
        bad, flagged, inset = oracle.check_period3(seed, first, 1 << 22, 20000)
        assert bad == 0
        assert 0.35 < flagged / inset < 0.46          # period-3 components: ~42 % of what is left
    bad, _, _ = oracle.check_period3(1337, 0, 1 << 22, 20000, limit=1.05)
    assert bad > 0


def test_sampler_prefilter_is_conservative(oracle):
    """The sampler retires candidates on an FP32 pre-classification (rejected / escapes at step 1 /
    at step 2) when the float value clears its threshold by a margin.  With the kernel's margins no
    decided sample may disagree with the reference's FP64 arithmetic (cudabrot.cu:284-298,
    :319-340); with the margins at zero the float error must show up, i.e. the check can fail."""
    for ship in (False, True):
        for seed, first in ((1337, 0), (7, 1 << 44)):
            bad, cnt = oracle.check_prefilter(seed, first, 1 << 24, ship=ship)
            assert bad == 0
            assert 0.15 < cnt[0] / (1 << 24) < 0.30         # what still needs the exact path
    bad, _ = oracle.check_prefilter(1337, 0, 1 << 26, m_rej=0.0, m_esc=0.0)
    assert bad > 0


def test_period4_component_test_is_conservative(oracle):
    """Same evidence for the period-4 test of the 80-register build: Newton for a small root of the
    multiplier cubic mu^3 - (3 - c^2) mu^2 + (3 + c^2 - c^3 - c^4) mu - (1 + 2c^2 + 3c^3 + 3c^4 + 3c^5
    + c^6), mu = lambda / 16, accepted when |mu| + 3 |p / p'| < 0.999 / 16.  Every flagged sample
    must run the reference's loop to max_iterations; a bound beyond |lambda| = 1 must be caught."""
    for seed, first in ((1337, 0), (99, 1 << 40)):
        bad, flagged, inset = oracle.check_period4(seed, first, 1 << 22, 20000)
        assert bad == 0
        assert 0.13 < flagged / inset < 0.21          # period-4 components: ~17.5 % of what is left
    bad, _, _ = oracle.check_period4(1337, 0, 1 << 22, 20000, mu_max=1.05 / 16)
    assert bad > 0


def test_cycle_certificate_is_conservative(oracle):
    """The kernel declares a sample `hit max` once it has found an attracting cycle of z^2 + c
    (period search in float, Newton in double, residual < 1e-12, |multiplier|^2 < 0.998): c then lies
    in a hyperbolic component.  Every certified sample must run the reference's loop
    (cudabrot.cu:319-340) to max_iterations; a multiplier bound beyond 1 must be caught by this very
    check.  Tried where the kernel tries it (from deep age 64) and from age 4, where it has to
    settle nearly everything the closed-form period-3/4 tests leave."""
    for seed, first in ((1337, 0), (99, 1 << 40)):
        bad, st = oracle.check_certificate(seed, first, 1 << 22, 20000)
        assert bad == 0
        assert st["certified"] > 0.3 * st["inset"]          # the rest turns bit-periodic before age 64
        assert st["iters_cert"] < 0.6 * st["iters_exact"]
        bad, st = oracle.check_certificate(seed, first, 1 << 22, 20000, first_age=4)
        assert bad == 0
        assert st["certified"] > 0.93 * st["inset"]
        assert st["iters_cert"] < 0.25 * st["iters_exact"]
    bad, _ = oracle.check_certificate(1337, 0, 1 << 22, 20000, lam2_max=1.1)
    assert bad > 0
    bad, _ = oracle.check_certificate(1337, 0, 1 << 22, 20000, lam2_max=1.1, first_age=4)
    assert bad > 0
