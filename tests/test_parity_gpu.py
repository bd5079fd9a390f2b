"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FULL = (-2.0, 2.0, -2.0, 2.0)

COUNTER_KEYS = ["candidates", "rejected", "hit_max", "too_early", "accepted", "escape_iters",
                "orbit_points", "increments"]


SHIP_CANVAS = (-2.2, 1.8, -2.5, 1.5)   # the ship sits below the real axis (imag grows downward)


def gpu_render(buddha, w, h, m, c, seed, first, count, canvas=FULL, flags=0):
    with buddha.Renderer(w, h, m, c, canvas=canvas, seed=seed, flags=flags) as r:
        r.render_samples(first, count)
        return r.read_histogram(), r.counters()


def assert_same(hist, cnt, ohist, ocnt):
    for k in COUNTER_KEYS:
        assert cnt[k] == ocnt[k], (k, cnt[k], ocnt[k])
    assert hist.dtype == np.uint32 and hist.shape == ohist.shape
    assert np.array_equal(hist, ohist)


# BASELINE.json configs at sizes the oracle finishes in seconds
CASES = {
    "cfg1_full_2^24": (1000, 1000, 100, 20, FULL, 1 << 24),
    "cfg2_deep": (4000, 4000, 20000, 10000, FULL, 1 << 22),
    "cfg3_small_canvas": (2000, 2000, 2000, 20, FULL, 1 << 22),
    "cfg4_zoom": (8000, 4000, 5000, 20, (0.0, 1.0, 0.0, 0.5), 1 << 21),
    "cfg5_ch1": (1000, 1000, 1000, 20, FULL, 1 << 21),
    "ragged_count": (333, 77, 300, 5, (-1.7, 0.3, -0.123, 0.777), 100003),
    "tiny": (3, 5, 64, 0, FULL, 4097),
    "min_gt_max": (64, 64, 50, 60, FULL, 1 << 16),
    "m_zero": (64, 64, 0, 0, FULL, 1 << 14),
    "m_one": (64, 64, 1, 0, FULL, 1 << 16),
    "m_33_block_tail": (128, 128, 37, 3, FULL, 1 << 18),
    "extreme_zoom_exact_binning": (500, 500, 400, 10,
                                   (-0.743644786, -0.7436447859, 0.1318252536, 0.1318252537),
                                   1 << 20),
}


@pytest.mark.parametrize("name", list(CASES))
def test_histogram_and_counters_match_oracle(buddha, oracle, name):
    w, h, m, c, canvas, n = CASES[name]
    ohist, ocnt, _ = oracle.render(w, h, m, c, 1337, 0, n, canvas=canvas)
    hist, cnt = gpu_render(buddha, w, h, m, c, 1337, 0, n, canvas)
    assert_same(hist, cnt, ohist, ocnt)


@pytest.mark.parametrize("m", [2, 3, 5, 6, 7, 13, 14, 15, 21, 22, 23, 24, 45, 46, 47, 48, 69,
                               70, 71, 94])
def test_tier_boundaries(buddha, oracle, m):
    """max-iter on, just below and just above every tier boundary of the escape test (2, 6, 22,
    then batches / rounds of 24: 46, 70, 94 ...), with cutoffs that accept escapes inside the
    first tiers."""
    for c in (0, 1, 2, 5, 6, 21, 22, 45):
        if c >= m:
            continue
        n = (1 << 17) + 77
        ohist, ocnt, _ = oracle.render(96, 64, m, c, 4242, 1 << 33, n)
        hist, cnt = gpu_render(buddha, 96, 64, m, c, 4242, 1 << 33, n)
        assert_same(hist, cnt, ohist, ocnt)


@pytest.mark.parametrize("flags_name", ["F_NO_SHORTCUT", "F_SIMPLE_KERNEL", "F_EXACT_BINNING"])
def test_kernel_variants_agree(buddha, oracle, flags_name):
    """The periodicity shortcut, the division-free binning and the persistent scheduling change
    nothing in the output."""
    flags = getattr(buddha, flags_name)
    for (w, h, m, c, canvas, n) in [(512, 512, 3000, 100, FULL, 1 << 20),
                                    (800, 400, 500, 20, (0.0, 1.0, 0.0, 0.5), 1 << 20)]:
        ohist, ocnt, _ = oracle.render(w, h, m, c, 7, 123456789, n, canvas=canvas)
        hist, cnt = gpu_render(buddha, w, h, m, c, 7, 123456789, n, canvas, flags=flags)
        assert_same(hist, cnt, ohist, ocnt)


@pytest.mark.parametrize("pool_mb", [None, "1"])
def test_tile_binned_scatter_agrees(buddha, oracle, monkeypatch, pool_mb):
    """The tile-binned scatter used for histograms far beyond L2 (forced here on small canvases:
    16 KB tiles, a 4 MB or 1 MB list pool so that lists overflow into the direct path) changes
    nothing, across the calibration launch, several render calls and list overflow."""
    if pool_mb:
        monkeypatch.setenv("BUDDHA_TILE_POOL_MB", pool_mb)
    for (w, h, m, c, canvas, n) in [(700, 500, 300, 10, FULL, (1 << 22) + 12345),
                                    (900, 300, 1000, 20, (0.0, 1.0, 0.0, 0.5), 1 << 21)]:
        ohist, ocnt, _ = oracle.render(w, h, m, c, 1337, 0, n, canvas=canvas)
        with buddha.Renderer(w, h, m, c, canvas=canvas, flags=buddha.F_FORCE_TILED) as r:
            r.render_samples(0, 1000)                 # shorter than the calibration window
            r.render_samples(1000, (1 << 20) - 1000)
            r.render_samples(1 << 20, n - (1 << 20))
            assert_same(r.read_histogram(), r.counters(), ohist, ocnt)


@pytest.mark.parametrize("extra", ["", "F_NO_SHORTCUT", "F_SIMPLE_KERNEL", "F_FORCE_TILED"])
def test_burning_ship_matches_oracle(buddha, oracle, extra):
    """BUDDHA_F_BURNING_SHIP = the reference built with RENDER_BURNING_SHIP (cudabrot.cu:15-17,
    :327-330, :353-356, :397-399)."""
    flags = buddha.F_BURNING_SHIP | (getattr(buddha, extra) if extra else 0)
    for (w, h, m, c, canvas, n) in [(600, 600, 500, 20, FULL, 1 << 20),
                                    (512, 384, 20000, 2000, SHIP_CANVAS, 1 << 19),
                                    (96, 64, 7, 0, FULL, 70001)]:
        ohist, ocnt, _ = oracle.render(w, h, m, c, 1337, 5, n, canvas=canvas, burning_ship=True)
        assert ocnt["rejected"] == 0
        hist, cnt = gpu_render(buddha, w, h, m, c, 1337, 5, n, canvas, flags=flags)
        assert_same(hist, cnt, ohist, ocnt)


FUSED_CASES = [
    # (w, h, canvas, channels [(max, min)], samples, flags)
    (500, 400, FULL, [(100, 20), (1000, 20), (20000, 20)], 1 << 21, ""),       # config 5's trio
    (333, 77, (-1.7, 0.3, -0.123, 0.777), [(3000, 50), (23, 0), (400, 399)], 300007, ""),
    (256, 256, FULL, [(60, 30), (2000, 1000)], 1 << 20, "F_NO_SHORTCUT"),      # disjoint windows
    (700, 500, FULL, [(300, 10), (5000, 100), (40, 5), (25, 24)], (1 << 21) + 5, "F_FORCE_TILED"),
    (300, 300, SHIP_CANVAS, [(200, 10), (5000, 200)], 1 << 19, "F_BURNING_SHIP"),
]


@pytest.mark.parametrize("case", range(len(FUSED_CASES)))
def test_fused_channels_match_separate_oracle_runs(buddha, oracle, case):
    """A fused multi-channel context renders every candidate once; each channel must be
    bit-identical -- histogram and counters -- to a separate single-channel render with that
    channel's (-m, -c) over the same sample indices (three runs of the reference in
    generate_hires_color_image.sh:27-59)."""
    w, h, canvas, channels, n, fl = FUSED_CASES[case]
    flags = getattr(buddha, fl) if fl else 0
    ship = fl == "F_BURNING_SHIP"
    first = 12345
    with buddha.Renderer(w, h, canvas=canvas, seed=99, flags=flags, channels=channels) as r:
        r.render_samples(first, 1000)
        r.render_samples(first + 1000, n - 1000)
        full = r.read_histogram()
        assert full.shape == (len(channels), h, w)
        for k, (m, c) in enumerate(channels):
            ohist, ocnt, _ = oracle.render(w, h, m, c, 99, first, n, canvas=canvas,
                                           burning_ship=ship)
            assert_same(r.read_channel(k), r.channel_counters(k), ohist, ocnt)
            assert np.array_equal(full[k], ohist)
            img, mx, scale = r.tonemap(2.2, big_endian=True, channel=k)
            oimg, omx, oscale = oracle.tonemap(ohist, 2.2, big_endian=True)
            assert (mx, scale) == (omx, oscale) and np.array_equal(img, oimg)


def test_cli_fused_channels(buddha, oracle, tmp_path):
    """--channels: one pass, one -s file and one PGM per channel, each byte-identical to what a
    single-channel run writes (generate_hires_color_image.sh:27-59 needs three runs)."""
    cli = buddha.capi.CLI_PATH
    save, out = str(tmp_path / "state.raw"), str(tmp_path / "img.pgm")
    channels = [(100, 20), (1000, 20), (5000, 200)]
    args = [cli, "-w", "240", "-h", "160", "-g", "2.2", "-s", save, "-o", out, "--samples",
            "400000", "--channels", ",".join("%d:%d" % mc for mc in channels)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    for k, (m, c) in enumerate(channels):
        ohist, _, _ = oracle.render(240, 160, m, c, 1337, 0, 400000)
        got = np.fromfile("%s.ch%d" % (save, k), dtype="<u4").reshape(160, 240)
        assert np.array_equal(got, ohist)
        oimg, _, _ = oracle.tonemap(ohist, 2.2)
        opgm = str(tmp_path / "oracle.pgm")
        oracle.write_pgm(opgm, oimg)
        assert open(str(tmp_path / ("img.ch%d.pgm" % k)), "rb").read() == open(opgm, "rb").read()
    # resume: the second run continues the stream and accumulates on top of the loaded channels
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Continuing the sample stream at index 400000." in r.stdout
    for k, (m, c) in enumerate(channels):
        ohist, _, _ = oracle.render(240, 160, m, c, 1337, 0, 800000)
        got = np.fromfile("%s.ch%d" % (save, k), dtype="<u4").reshape(160, 240)
        assert np.array_equal(got, ohist)


def test_fused_rejects_bad_channels(buddha):
    with pytest.raises(buddha.capi.BuddhaError):
        buddha.Renderer(64, 64, channels=[(100, 20), (22, 0)])       # max must exceed the tiers
    with pytest.raises(buddha.capi.BuddhaError):
        buddha.Renderer(64, 64, channels=[(100, 20)] * 5)


def test_shortcut_statistics(buddha):
    with buddha.Renderer(256, 256, 20000, 10000) as r:
        r.render_samples(0, 1 << 22)
        c = r.counters()
    assert c["shortcut_hits"] > 0.5 * c["hit_max"]
    assert c["executed_iters"] < c["escape_iters"]
    with buddha.Renderer(256, 256, 20000, 10000, flags=buddha.F_NO_SHORTCUT) as r:
        r.render_samples(0, 1 << 20)
        c = r.counters()
    assert c["shortcut_hits"] == 0 and c["executed_iters"] >= c["escape_iters"]


def test_sample_ranges_are_additive_and_64bit(buddha, oracle):
    """Any split of an index range over calls gives the same summed histogram; indices beyond
    2^32 use the high counter word."""
    first = (1 << 40) + 12345
    n = 300000
    ohist, ocnt, _ = oracle.render(200, 200, 150, 10, 99, first, n)
    with buddha.Renderer(200, 200, 150, 10, seed=99) as r:
        r.render_samples(first, 1000)
        r.render_samples(first + 1000, 0)
        r.render_samples(first + 1000, 4096 * 3 + 5)
        r.render_samples_async(first + 1000 + 4096 * 3 + 5, n - (1000 + 4096 * 3 + 5))
        r.sync()
        assert r.last_render_ms() > 0
        assert_same(r.read_histogram(), r.counters(), ohist, ocnt)
        r.reset_counters()
        assert r.counters()["candidates"] == 0


def test_full_size_cfg3_histogram(buddha, oracle):
    """20000x20000 (1.6 GB histogram, BASELINE config 3) at a reduced sample count."""
    n = 1 << 21
    ohist, ocnt, _ = oracle.render(20000, 20000, 2000, 20, 1337, 0, n)
    hist, cnt = gpu_render(buddha, 20000, 20000, 2000, 20, 1337, 0, n)
    assert_same(hist, cnt, ohist, ocnt)
    assert int(hist.sum(dtype=np.uint64)) == cnt["increments"]


def test_full_size_properties_cfg2(buddha, monkeypatch):
    """BASELINE config 2 at its full canvas and 2^30 samples, where the oracle would need minutes:
    size-independent properties instead.  One call == any split into calls; the histogram sum ==
    the increments counted; every candidate ends in exactly one class; the exact shortcut and the
    division-free binning change nothing (checked on a 2^27-sample prefix), nor does the size of the
    certificate queues."""
    w = h = 4000
    n = 1 << 30
    with buddha.Renderer(w, h, 20000, 10000) as r:
        r.render_samples(0, n)
        whole, cnt = r.read_histogram(), r.counters()
    assert int(whole.sum(dtype=np.uint64)) == cnt["increments"]
    assert cnt["candidates"] == n == (cnt["rejected"] + cnt["hit_max"] + cnt["too_early"] +
                                      cnt["accepted"])
    assert cnt["orbit_points"] >= cnt["increments"] > 0
    assert cnt["orbit_points"] >= 10001 * cnt["accepted"]         # every accepted i >= min
    with buddha.Renderer(w, h, 20000, 10000) as r:
        cuts = [0, 12345, 4096 * 1000 + 7, 1 << 29, n]
        for a, b in zip(cuts, cuts[1:]):
            r.render_samples(a, b - a)
        parts, cnt2 = r.read_histogram(), r.counters()
    assert np.array_equal(whole, parts)
    for k in COUNTER_KEYS:
        assert cnt[k] == cnt2[k], k
    m = 1 << 27
    ref = None
    for flags in (0, buddha.F_NO_SHORTCUT, buddha.F_EXACT_BINNING):
        with buddha.Renderer(w, h, 20000, 10000, flags=flags) as r:
            r.render_samples(1 << 40, m)
            hist, c = r.read_histogram(), r.counters()
        if ref is None:
            ref = (hist, c)
        else:
            assert np.array_equal(ref[0], hist)
            for k in COUNTER_KEYS:
                assert ref[1][k] == c[k], k
    # the cycle certificate with its smallest queue (160 entries per warp): at this sample count
    # the queues fill and are worked off several times per warp (with the default 512 only the
    # single 2^30-sample call above gets there)
    monkeypatch.setenv("BUDDHA_CERT_QUEUE", "160")
    with buddha.Renderer(w, h, 20000, 10000) as r:
        r.render_samples(1 << 40, m)
        hist, c = r.read_histogram(), r.counters()
    assert np.array_equal(ref[0], hist)
    for k in COUNTER_KEYS:
        assert ref[1][k] == c[k], k
    assert c["executed_iters"] < ref[1]["executed_iters"]      # (ref[1]: the run with every shortcut)


def test_full_size_properties_cfg3_tiled_vs_direct(buddha, monkeypatch):
    """BASELINE config 3 at its full 20000x20000 canvas: the tile-binned scatter (default for
    1.6 GB) and the direct reductions must produce the same histogram over 2^28 samples, across the
    calibration launch and a multi-launch pipeline."""
    n = (1 << 28) + 4097
    with buddha.Renderer(20000, 20000, 2000, 20) as r:
        r.render_samples(0, n)
        tiled, cnt = r.read_histogram(), r.counters()
    assert int(tiled.sum(dtype=np.uint64)) == cnt["increments"]
    monkeypatch.setenv("BUDDHA_TILE_MIN_MB", "1000000")
    with buddha.Renderer(20000, 20000, 2000, 20) as r:
        r.render_samples(0, 1 << 27)
        r.render_samples(1 << 27, n - (1 << 27))
        direct, cnt2 = r.read_histogram(), r.counters()
    assert np.array_equal(tiled, direct)
    for k in COUNTER_KEYS:
        assert cnt[k] == cnt2[k], k


def test_load_accumulate_read_roundtrip(buddha, oracle):
    """-s semantics (cudabrot.cu:215-280): a loaded buffer is accumulated into, not replaced."""
    rng = np.random.default_rng(5)
    base = rng.integers(0, 1000, size=(120, 160), dtype=np.uint32)
    ohist, _, _ = oracle.render(160, 120, 100, 20, 1337, 0, 1 << 18, hist=base.copy())
    with buddha.Renderer(160, 120, 100, 20) as r:
        r.load_histogram(base)
        assert np.array_equal(r.read_histogram(), base)
        r.render_samples(0, 1 << 18)
        assert np.array_equal(r.read_histogram(), ohist)
        r.clear()
        assert r.read_histogram().sum() == 0
        with pytest.raises(buddha.BuddhaError):
            r.load_histogram(np.zeros(7, dtype=np.uint32))


def test_render_seconds_semantics(buddha):
    """-t 0 renders exactly one pass of 512*512*50 candidates (cudabrot.cu:488-491, quirk 6)."""
    with buddha.Renderer(256, 256, 100, 20) as r:
        done, passes = r.render_seconds(0.0)
        assert (done, passes) == (13107200, 1)
        assert r.counters()["candidates"] == 13107200
        done2, passes2 = r.render_seconds(0.3, first=done)
        assert passes2 >= 2 and done2 > 13107200


@pytest.mark.parametrize("gamma", [1.0, 2.2, 0.5, -1.0])
@pytest.mark.parametrize("big_endian", [False, True])
def test_tonemap_matches_oracle_on_render(buddha, oracle, gamma, big_endian):
    with buddha.Renderer(640, 480, 200, 20) as r:
        r.render_samples(0, 1 << 22)
        hist = r.read_histogram()
        img, mx, scale = r.tonemap(gamma, big_endian)
        assert r.last_tonemap_ms() > 0
    oimg, omx, oscale = oracle.tonemap(hist, gamma, big_endian)
    assert (mx, scale) == (omx, oscale)
    assert np.array_equal(img, oimg)


@pytest.mark.parametrize("mx", [5, 70000, (1 << 18) - 1, 1 << 18, 3000000, 102438, (1 << 32) - 1])
@pytest.mark.parametrize("gamma", [1.0, 2.2, 0.37, -1.0])
def test_tonemap_table_and_threshold_paths(buddha, oracle, mx, gamma):
    """Counts below 2^18 go through a full table, larger ones through the threshold search (whose
    thresholds come from a closed-form guess made exact by a short walk): both must reproduce the
    reference expression (glibc pow included) for every count, on either side of the switch."""
    rng = np.random.default_rng(mx % 1000 + 7)
    hist = rng.integers(0, mx, size=(96, 128), dtype=np.uint64, endpoint=True).astype(np.uint32)
    # dense low counts, the region around the table limit and the maximum itself
    hist[0, :] = np.minimum(np.arange(128, dtype=np.uint64), mx).astype(np.uint32)
    hist[1, :] = np.minimum((1 << 18) - 64 + np.arange(128, dtype=np.uint64), mx).astype(np.uint32)
    hist[2, :] = np.uint32(mx) - np.minimum(np.arange(128, dtype=np.uint64), mx).astype(np.uint32)
    with buddha.Renderer(128, 96, 10, 0) as r:
        r.load_histogram(hist)
        for be in (False, True):
            img, m, scale = r.tonemap(gamma, big_endian=be)
            oimg, om, oscale = oracle.tonemap(hist, gamma, big_endian=be)
            assert (m, scale) == (om, oscale) == (mx, 65535.0 / mx)
            assert np.array_equal(img, oimg)


@pytest.mark.parametrize("name,side", [("kat64", (64, 64)), ("big", (32, 32)), ("zero", (16, 16))])
@pytest.mark.parametrize("gamma", ["1.0", "2.2", "0.5", "-1"])
def test_tonemap_matches_reference_golden_pgm(buddha, tmp_path, name, side, gamma):
    """Against files written by the reference's own host code (tests/golden/make_golden.py);
    `big` has counts past the 2^22-entry table, i.e. the threshold-search path."""
    hist = np.fromfile(os.path.join(GOLDEN, "hist_%s.raw" % name), dtype="<u4").reshape(side)
    with buddha.Renderer(side[1], side[0], 10, 0) as r:
        r.load_histogram(hist)
        img, mx, scale = r.tonemap(float(gamma), big_endian=True)
    out = str(tmp_path / "o.pgm")
    buddha.write_pgm(out, img, side[1], side[0])
    golden = os.path.join(GOLDEN, "tonemap_%s_g%s.pgm" % (name, gamma))
    assert open(out, "rb").read() == open(golden, "rb").read()
    assert open(golden + ".stdout").read().strip() == "Max value: %d, scale: %f" % (mx, scale)


def test_reference_device_code_agrees_with_oracle(oracle, tmp_path):
    """Pins the oracle to the REFERENCE's real SASS: oracle/_ref/ref_probe runs the reference's
    own InMainCardioid / InOrder2Bulb / IterateMandelbrot / IterateAndRecord (cudabrot.cu:284-365)
    on this GPU over the oracle's Philox sample list."""
    if not os.path.exists(oracle.REF_PROBE):
        pytest.fail("oracle/_ref/ref_probe missing: run make -C oracle where /root/reference exists")
    for (w, h, canvas, m, c, first, n) in [(1000, 1000, FULL, 100, 20, 0, 1 << 20),
                                           (200, 100, (0.0, 1.0, 0.0, 0.5), 1000, 20, 0, 1 << 18),
                                           (333, 77, (-1.7, 0.3, -0.123, 0.777), 3000, 50, 1 << 33,
                                            1 << 17)]:
        pts = np.empty((n, 2), dtype=np.float64)
        L = oracle.lib()
        import ctypes as C
        re, im = C.c_double(), C.c_double()
        for k in range(n):
            L.oracle_sample(1337, first + k, C.byref(re), C.byref(im))
            pts[k, 0], pts[k, 1] = re.value, im.value
        sp, ip, hp = (str(tmp_path / f) for f in ("s.f64", "i.i32", "h.raw"))
        pts.tofile(sp)
        r = oracle.run_ref_probe("orbits", w, h, repr(canvas[0]), repr(canvas[1]), repr(canvas[2]),
                                 repr(canvas[3]), m, c, sp, ip, hp)
        assert r.returncode == 0, r.stdout + r.stderr
        ref_iters = np.fromfile(ip, dtype=np.int32)
        ref_hist = np.fromfile(hp, dtype=np.uint32).reshape(h, w)
        assert np.array_equal(ref_iters, oracle.classify(1337, first, n, m))
        ohist, _, _ = oracle.render(w, h, m, c, 1337, first, n, canvas=canvas)
        assert np.array_equal(ref_hist, ohist)
        # the same harness compiled with -DRENDER_BURNING_SHIP pins the ship variant
        if not os.path.exists(oracle.REF_PROBE_SHIP):
            pytest.fail("oracle/_ref/ref_probe_ship missing")
        r = oracle.run_ref_probe("orbits", w, h, repr(canvas[0]), repr(canvas[1]), repr(canvas[2]),
                                 repr(canvas[3]), m, c, sp, ip, hp, burning_ship=True)
        assert r.returncode == 0, r.stdout + r.stderr
        ref_iters = np.fromfile(ip, dtype=np.int32)
        ref_hist = np.fromfile(hp, dtype=np.uint32).reshape(h, w)
        assert np.array_equal(ref_iters, oracle.classify(1337, first, n, m, burning_ship=True))
        ohist, _, _ = oracle.render(w, h, m, c, 1337, first, n, canvas=canvas, burning_ship=True)
        assert np.array_equal(ref_hist, ohist)


def test_cli_end_to_end(buddha, oracle, tmp_path):
    """The drop-in binary: --samples render, -s file layout, PGM bytes, resume accumulation."""
    cli = buddha.capi.CLI_PATH
    save, out = str(tmp_path / "state.raw"), str(tmp_path / "img.pgm")
    args = [cli, "-w", "300", "-h", "200", "-m", "200", "-c", "10", "-g", "2.2", "-s", save,
            "-o", out, "--samples", "500000"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Creating 300x200 image, 200 max iterations." in r.stdout
    assert "File %s doesn't exist yet. Not loading." % save in r.stdout
    assert "Done! Output image saved: %s" % out in r.stdout
    ohist, _, _ = oracle.render(300, 200, 200, 10, 1337, 0, 500000)
    assert np.array_equal(np.fromfile(save, dtype="<u4").reshape(200, 300), ohist)
    oimg, omx, oscale = oracle.tonemap(ohist, 2.2)
    opgm = str(tmp_path / "oracle.pgm")
    oracle.write_pgm(opgm, oimg)
    assert open(out, "rb").read() == open(opgm, "rb").read()
    assert "Max value: %d, scale: %f" % (omx, oscale) in r.stdout
    # second run resumes: loads the buffer and continues the Philox stream at the sidecar cursor
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Continuing the sample stream at index 500000." in r.stdout
    ohist2, _, _ = oracle.render(300, 200, 200, 10, 1337, 500000, 500000, hist=ohist.copy())
    assert np.array_equal(np.fromfile(save, dtype="<u4").reshape(200, 300), ohist2)
    # wrong-size -s file is refused with the reference's message and exit code 1 (:239-245)
    r = subprocess.run([cli, "-w", "301", "-h", "200", "-s", save, "--samples", "10"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 1
    assert "doesn't match the expected size of %d bytes." % (301 * 200 * 4) in r.stdout
    # -t 0 = exactly one reference-sized pass
    r = subprocess.run([cli, "-w", "100", "-h", "100", "-t", "0", "-o", out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "1 Buddhabrot passes took" in r.stdout
    assert "13107200 candidate samples" in r.stdout


def test_in_process_multi_gpu_merge(buddha, oracle):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    n = 1 << 20
    ohist, _, _ = oracle.render(400, 300, 300, 20, 1337, 0, n)
    rs = [buddha.Renderer(400, 300, 300, 20, device=d) for d in range(2)]
    rs[0].render_samples(0, n // 2)
    rs[1].render_samples(n // 2, n - n // 2)
    buddha.merge_in_process(rs, root=0)
    assert np.array_equal(rs[0].read_histogram(), ohist)
    for r in rs:
        r.close()


def test_in_process_fused_merge_and_tonemap_on_other_device(buddha, oracle):
    """A fused context on device 1 must tone-map there even while device 0 is current, and
    buddha_merge must leave the caller's device alone (round-1 review: the channel assembly ran
    before cudaSetDevice).  Channels are compared with separate oracle runs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    chans = [(100, 20), (1000, 20)]
    n = 1 << 19
    rs = [buddha.Renderer(300, 200, channels=chans, device=d) for d in (1, 0)]   # root on device 1
    rs[0].render_samples(0, n // 2)
    rs[1].render_samples(n // 2, n - n // 2)
    torch.cuda.set_device(0)
    buddha.merge_in_process(rs, root=0)
    assert torch.cuda.current_device() == 0
    for k, (m, c) in enumerate(chans):
        ohist, _, _ = oracle.render(300, 200, m, c, 1337, 0, n)
        oimg, omx, oscale = oracle.tonemap(ohist, 2.2, big_endian=True)
        img, mx, scale = rs[0].tonemap(2.2, big_endian=True, channel=k)   # device 0 is current
        assert (mx, scale) == (omx, oscale) and np.array_equal(img, oimg)
        assert np.array_equal(rs[0].read_channel(k), ohist)
    # contexts that do not describe the same canvas / channels are refused
    other = buddha.Renderer(300, 200, channels=[(100, 20), (999, 20)], device=0)
    with pytest.raises(buddha.BuddhaError):
        buddha.merge_in_process([rs[0], other], root=0)
    for r in rs + [other]:
        r.close()


def test_histogram_digest_matches_numpy_restatement(buddha, oracle):
    """buddha_histogram_digest (blocked FNV-1a-64 formed on the GPU) == the numpy restatement over
    the oracle's histogram; ragged sizes, a fused channel, and sensitivity to a single cell."""
    for (w, h, m, c, n) in [(1000, 1000, 100, 20, 1 << 22), (333, 77, 300, 5, 100003),
                            (64, 64, 50, 60, 1 << 12)]:
        ohist, _, _ = oracle.render(w, h, m, c, 1337, 0, n)
        with buddha.Renderer(w, h, m, c) as r:
            assert r.digest() == oracle.blocked_fnv(np.zeros((h, w), dtype=np.uint32))
            r.render_samples(0, n)
            assert r.digest() == oracle.blocked_fnv(ohist)
            bumped = ohist.copy()
            bumped[h - 1, w - 1] += 1
            r.load_histogram(bumped)
            assert r.digest() == oracle.blocked_fnv(bumped) != oracle.blocked_fnv(ohist)
    chans = [(100, 20), (1000, 20), (5000, 20)]
    with buddha.Renderer(500, 400, channels=chans) as r:
        r.render_samples(0, 1 << 20)
        for k, (m, c) in enumerate(chans):
            ohist, _, _ = oracle.render(500, 400, m, c, 1337, 0, 1 << 20)
            assert r.digest(k) == oracle.blocked_fnv(ohist)


@pytest.mark.parametrize("fused", [False, True])
def test_overlapped_transfers_match_blocking_calls(buddha, oracle, fused):
    """add_histogram_async / snapshot / read_snapshot / tonemap_snapshot (copies overlapped with
    the next render) deliver exactly what load / read / tonemap deliver at the same point."""
    import torch
    w, h, n = 640, 480, 1 << 20
    chans = [(100, 20), (1000, 20)] if fused else None
    n_ch = 2 if fused else 1
    rng = np.random.default_rng(11)
    saved = torch.zeros(n_ch * h * w, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    saved[:] = rng.integers(0, 50, size=saved.size, dtype=np.uint32)
    shape = (n_ch, h, w) if fused else (h, w)
    expect = []
    for (m, c) in (chans or [(300, 20)]):
        base = saved.reshape(n_ch, h, w)[len(expect)].copy()
        h1, _, _ = oracle.render(w, h, m, c, 1337, 0, n, hist=base.copy())
        h2, _, _ = oracle.render(w, h, m, c, 1337, n, n, hist=h1 + base)   # saved counts added twice
        expect.append((h1, h2))
    with buddha.Renderer(w, h, 300, 20, channels=chans) as r:
        with pytest.raises(buddha.BuddhaError):
            r.read_snapshot()                       # no snapshot yet
        r.add_histogram_async(saved.reshape(shape))
        r.render_samples_async(0, n)
        r.snapshot()
        r.add_histogram_async(saved.reshape(shape))  # both enqueued behind the snapshot
        r.render_samples_async(n, n)
        snap = r.read_snapshot()                     # step 1, while step 2 renders
        for k in range(n_ch):
            got = snap[k] if fused else snap
            assert np.array_equal(got, expect[k][0])
            oimg, omx, oscale = oracle.tonemap(expect[k][0], 2.2, big_endian=True)
            img, mx, scale = r.tonemap_snapshot(2.2, big_endian=True, channel=k)
            assert (mx, scale) == (omx, oscale) and np.array_equal(img, oimg)
        r.sync()
        live = r.read_histogram()
        for k in range(n_ch):
            assert np.array_equal(live[k] if fused else live, expect[k][1])
        with pytest.raises(buddha.BuddhaError):
            r.add_histogram_async(np.zeros(7, dtype=np.uint32))


def test_stop_flag_ends_an_endless_run(buddha, oracle):
    """-t -1 semantics (cudabrot.cu:483, :756-760): the run ends at the next pass boundary once
    the flag is set, and what was rendered is exactly the index range [0, samples_done)."""
    import ctypes
    import threading
    with buddha.Renderer(200, 150, 200, 20) as r:
        with pytest.raises(buddha.BuddhaError):
            r.render_seconds(-1.0)                   # endless without a stop flag is refused
        stop = ctypes.c_int(0)
        threading.Timer(0.4, lambda: setattr(stop, "value", 1)).start()
        done, passes = r.render_seconds(-1.0, stop=stop)
        assert passes >= 1 and done >= 13107200
        hist, cnt = r.read_histogram(), r.counters()
        assert cnt["candidates"] == done
        assert int(hist.sum(dtype=np.uint64)) == cnt["increments"]
    with buddha.Renderer(200, 150, 200, 20) as r:   # the same range in one call: same histogram
        r.render_samples(0, done)
        assert np.array_equal(r.read_histogram(), hist)
    n = 1 << 22                                      # and its prefix agrees with the oracle
    ohist, ocnt, _ = oracle.render(200, 150, 200, 20, 1337, 0, n)
    with buddha.Renderer(200, 150, 200, 20) as r:
        r.render_samples(0, n)
        assert_same(r.read_histogram(), r.counters(), ohist, ocnt)


def test_cli_sigint_saves_state_and_exits_zero(buddha, oracle, tmp_path):
    """bin/cudabrot -t -1: SIGINT ends the run after the current pass, the -s buffer and the PGM
    are written, the exit code is 0 (cudabrot.cu:483, :756-760, :785), and the buffer equals the
    oracle over [0, cursor)."""
    import signal
    import time
    cli = buddha.capi.CLI_PATH
    save, out = str(tmp_path / "state.raw"), str(tmp_path / "img.pgm")
    p = subprocess.Popen([cli, "-w", "160", "-h", "120", "-m", "60", "-c", "5", "-t", "-1", "-s",
                          save, "-o", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True)
    time.sleep(2.5)
    p.send_signal(signal.SIGINT)
    stdout, _ = p.communicate(timeout=120)
    assert p.returncode == 0, stdout
    assert "Press ctrl+C to finish." in stdout and "Signal 2 received" in stdout
    assert "Done! Output image saved: %s" % out in stdout
    seed, nxt = open(save + ".cursor").read().split()[1::2]
    nxt = int(nxt)
    assert int(seed) == 1337 and nxt >= 13107200
    got = np.fromfile(save, dtype="<u4").reshape(120, 160)
    m = 1 << 23   # the oracle over the whole range would take minutes: a prefix bounds it from below,
    ohist, _, _ = oracle.render(160, 120, 60, 5, 1337, 0, m)      # the library re-renders the range
    assert np.all(got >= ohist)
    with buddha.Renderer(160, 120, 60, 5) as r:
        r.render_samples(0, nxt)
        assert np.array_equal(r.read_histogram(), got)
        r.clear()
        r.render_samples(0, m)
        assert np.array_equal(r.read_histogram(), ohist)


def test_cli_refuses_partial_channel_set(buddha, tmp_path):
    """A fused -s resume with only some of the channel files present must stop instead of
    restarting the stream and overwriting the files that exist."""
    cli = buddha.capi.CLI_PATH
    save, out = str(tmp_path / "state.raw"), str(tmp_path / "img.pgm")
    args = [cli, "-w", "64", "-h", "48", "--channels", "100:20,1000:20", "-s", save, "-o", out,
            "--samples", "100000"]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    before = open(save + ".ch0", "rb").read()
    os.remove(save + ".ch1")
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 1
    assert "Only 1 of the 2 channel files" in r.stdout
    assert open(save + ".ch0", "rb").read() == before


def test_production_tiles_cfg5_fused_and_cfg3_m20000(buddha, oracle):
    """Oracle parity at the production tile size (64 MB tiles, calibration + pipeline launches):
    BASELINE config 5 fused at 10000x10000 and config 3's canvas at -m 20000."""
    n = 1 << 21
    chans = [(100, 20), (1000, 20), (20000, 20)]
    with buddha.Renderer(10000, 10000, channels=chans) as r:
        r.render_samples(0, n)
        for k, (m, c) in enumerate(chans):
            ohist, ocnt, _ = oracle.render(10000, 10000, m, c, 1337, 0, n)
            assert r.digest(k) == oracle.blocked_fnv(ohist)
            assert np.array_equal(r.read_channel(k), ohist)
            cc = r.channel_counters(k)
            for key in COUNTER_KEYS:
                assert cc[key] == ocnt[key], (k, key)
    ohist, ocnt, _ = oracle.render(20000, 20000, 20000, 20, 1337, 0, n)
    with buddha.Renderer(20000, 20000, 20000, 20) as r:
        r.render_samples(0, 1 << 20)            # calibration launch + a second, pipelined call
        r.render_samples(1 << 20, n - (1 << 20))
        assert r.digest() == oracle.blocked_fnv(ohist)
        assert_same(r.read_histogram(), r.counters(), ohist, ocnt)


def test_colour_combine_matches_numpy_restatement(buddha, oracle):
    """buddha_combine_rgb_u16 (the step generate_hires_color_image.sh:61-71 leaves to external
    tools): RGB = the three tone-mapped channels, bit for bit; HSL = the standard HSL -> RGB in
    double precision on the oracle's tone-mapped planes (tolerance: 1 of 65535, the rounding of
    the last multiplication may differ between libm-free device code and numpy)."""
    chans = [(100, 20), (1000, 20), (5000, 20)]
    w, h, n = 400, 300, 1 << 21
    planes = []
    for (m, c) in chans:
        ohist, _, _ = oracle.render(w, h, m, c, 1337, 0, n)
        img, mx, _ = oracle.tonemap(ohist, 2.2)
        planes.append((img.astype(np.float64), mx, img))
    with buddha.Renderer(w, h, channels=chans) as r:
        r.render_samples(0, n)
        rgb, mx = r.combine_rgb((2, 1, 0), gamma=2.2, mode="rgb")
        assert mx == [planes[2][1], planes[1][1], planes[0][1]]
        for k, ch in enumerate((2, 1, 0)):
            assert np.array_equal(rgb[:, :, k], planes[ch][2])
        be, _ = r.combine_rgb((2, 1, 0), gamma=2.2, mode="rgb", big_endian=True)
        assert np.array_equal(be.byteswap(), rgb)
        hsl, _ = r.combine_rgb((1, 0, 2), gamma=2.2, mode="hsl", hue_adjust=0.3)
        with pytest.raises(buddha.BuddhaError):
            r.combine_rgb((0, 1, 7))
    hue = (planes[1][0] / 65535.0 + 0.3) % 1.0
    sat, lig = planes[0][0] / 65535.0, planes[2][0] / 65535.0
    chroma = (1.0 - np.abs(2.0 * lig - 1.0)) * sat
    h6 = hue * 6.0
    x = chroma * (1.0 - np.abs(h6 % 2.0 - 1.0))
    sector = h6.astype(np.int64)
    zero = np.zeros_like(chroma)
    r1 = np.choose(sector, [chroma, x, zero, zero, x, chroma])
    g1 = np.choose(sector, [x, chroma, chroma, x, zero, zero])
    b1 = np.choose(sector, [zero, zero, x, chroma, chroma, x])
    m = lig - chroma * 0.5
    want = np.stack([np.rint(np.clip((v + m) * 65535.0, 0, 65535)) for v in (r1, g1, b1)], axis=2)
    diff = np.abs(hsl.astype(np.int64) - want.astype(np.int64))
    assert diff.max() <= 1
    assert (diff == 0).mean() > 0.999


def test_cli_colour_ppm(buddha, tmp_path):
    """--color rgb: the PPM is the three channel PGMs interleaved (P6, 16-bit big-endian)."""
    cli = buddha.capi.CLI_PATH
    out = str(tmp_path / "img.pgm")
    r = subprocess.run([cli, "-w", "120", "-h", "80", "--channels", "100:20,1000:20,5000:20",
                        "--color", "rgb", "-g", "2.2", "-o", out, "--samples", "300000"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    ppm = open(str(tmp_path / "img.color.ppm"), "rb").read()
    head = b"P6\n120 80\n65535\n"
    assert ppm.startswith(head)
    rgb = np.frombuffer(ppm[len(head):], dtype=">u2").reshape(80, 120, 3)
    for k in range(3):
        pgm = open(str(tmp_path / ("img.ch%d.pgm" % k)), "rb").read()
        grey = np.frombuffer(pgm[len(b"P5\n120 80\n65535\n"):], dtype=">u2").reshape(80, 120)
        assert np.array_equal(rgb[:, :, k], grey)
    r = subprocess.run([cli, "--color", "rgb", "--samples", "10"], capture_output=True, text=True,
                       timeout=300)
    assert r.returncode == 0 and "--color needs --channels" in r.stdout
