#!/usr/bin/env python
"""bench.py -- candidate samples/s (and orbit points/s) of the Buddhabrot hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]
                    [--workload cfg2|cfg1|cfg3|cfg3_m20000|cfg4|cfg5|cfg5a|cfg5b|cfg5c|cfg5_4k...]

A "step" renders one batch of candidate samples (a fresh block of Philox sample indices per step
and per rank) into the resident histogram.  N>1: one process per GPU (torchrun), disjoint sample
ranges, no data-path collective per step, one reduce(sum) of the histograms at the end of the timed
region (weak scaling).  Rank 0 prints ONE JSON line.

--impl reference times the UNMODIFIED reference program (oracle/_ref/cudabrot_ref = cudabrot.cu
compiled for sm_100a with only the arch flags changed) on one B200 -- cudabrot has no CPU path, and
BASELINE.json's north_star names exactly this binary as the ">= 3x" denominator -- and falls back
to the CPU oracle (OpenMP) only if that binary is missing.  The CPU oracle is always timed as
`cpu_baseline` (rank 0, N=1 of the native arm).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

# stdout carries exactly one JSON line: NCCL's own debug output goes to stderr, and so does
# anything else a library writes to file descriptor 1 (see claim_stdout)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

FULL = (-2.0, 2.0, -2.0, 2.0)
# name -> (w, h, max_iter, min_iter, canvas, default samples per step per GPU) ; BASELINE.json configs
WORKLOADS = {
    "cfg1": (1000, 1000, 100, 20, FULL, 1 << 32),
    "cfg2": (4000, 4000, 20000, 10000, FULL, 1 << 35),
    # 1.6 GB in + 2.4 GB out per step over PCIe take ~140 ms, and the last step's read-back has
    # nothing to hide behind: 2^35-sample steps (~420 ms of rendering) for the overlapped pipeline
    "cfg3": (20000, 20000, 2000, 20, FULL, 1 << 35),
    "cfg3_m20000": (20000, 20000, 20000, 20, FULL, 1 << 32),
    "cfg4": (8000, 4000, 5000, 20, (0.0, 1.0, 0.0, 0.5), 1 << 32),
    "cfg5a": (10000, 10000, 100, 20, FULL, 1 << 32),
    "cfg5b": (10000, 10000, 1000, 20, FULL, 1 << 32),
    "cfg5c": (10000, 10000, 20000, 20, FULL, 1 << 32),
    # config 5 as ONE fused pass: every candidate is rendered once into the three channels
    "cfg5": (10000, 10000, 20000, 20, FULL, 1 << 33),
    # the same trio on a canvas whose three histograms fit L2 together (192 MB), fused and apart
    "cfg5_4k": (4000, 4000, 20000, 20, FULL, 1 << 32),
    "cfg5_4k_a": (4000, 4000, 100, 20, FULL, 1 << 32),
    "cfg5_4k_b": (4000, 4000, 1000, 20, FULL, 1 << 32),
    "cfg5_4k_c": (4000, 4000, 20000, 20, FULL, 1 << 32),
    # dense canvases: every orbit point lands in a few cells (hot-address reductions); the question
    # north_star (d) raises -- would shared-memory privatised tiles pay? -- is answered by whether
    # the rate drops as the canvas shrinks (profiles/r02_dense_canvas.txt)
    "dense64": (64, 64, 1000, 20, FULL, 1 << 32),
    "dense256": (256, 256, 1000, 20, FULL, 1 << 32),
    "dense1k": (1000, 1000, 1000, 20, FULL, 1 << 32),
    "dense4k": (4000, 4000, 1000, 20, FULL, 1 << 32),
    # a zoom that keeps most of the orbit points: the 2000x2000 window around the main body
    "zoom_dense": (2000, 2000, 1000, 20, (-1.6, 0.6, -1.1, 1.1), 1 << 32),
}
# fused multi-channel workloads: [(max-iter, min-cutoff)] per channel (BASELINE.json configs[4])
CHANNELS = {"cfg5": [(100, 20), (1000, 20), (20000, 20)],
            "cfg5_4k": [(100, 20), (1000, 20), (20000, 20)]}
REFERENCE_PASS = 13107200  # 512 blocks * 512 threads * 50 samples, cudabrot.cu:20,23,34
# dram__bytes_read.sum + dram__bytes_write.sum of ONE render_persistent_kernel launch over 2^30
# samples, from the `ncu --set full` captures summarised in profiles/r02_summary.md
# (config 2: 40 KB before the cycle certificate; its queues -- 48 B per parked sample, 0.17 % of the
#  candidates -- are the 80 MB now: a quarter of the 2^32-sample capture r02_render_cfg2_cert_*)
NCU_DRAM_BYTES_PER_2P30_LAUNCH = {"cfg1": 4038912 + 5376, "cfg2": (16600320 + 304296704) // 4,
                                  "cfg3": 255045120 + 5058764000, "cfg4": 127435520 + 2082458000}
METRIC = "candidate samples/sec"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


_json_out = None


def claim_stdout():
    """Keep the real stdout for the one JSON line and point file descriptor 1 at stderr, so that C
    libraries printing to stdout (NCCL's version banner under NCCL_DEBUG=VERSION) cannot get in
    front of it."""
    global _json_out
    if _json_out is None:
        sys.stdout.flush()
        _json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _json_out or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms; only samples taken between
    mark_start() and stop() (the timed region) are reported."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.rows, self.proc, self.device, self.t_start = [], None, device, None

    def start(self):
        """Launch the sampler (nvidia-smi takes a while to produce its first line: call this well
        before the timed region)."""
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        """Block until nvidia-smi has produced its first line (it can take more than a second on a
        fresh box), so that the timed region is sampled from its start."""
        t0 = time.perf_counter()
        while self.proc and not self.rows and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark_start(self):
        self.t_start = time.perf_counter()

    def stop(self):
        t_end = time.perf_counter()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, pw, reasons = [], 0, [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if self.t_start is not None and not (self.t_start <= t <= t_end + 0.05):
                continue
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1])); pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def time_cpu_oracle(wl, budget_s=12.0, channels=None):
    """cpu_baseline: the OpenMP oracle on this host's cores, on a bounded sample of the workload.
    Multi-channel workloads: one oracle run per channel, as the reference is used
    (generate_hires_color_image.sh:27-59); the rate is samples / total time."""
    import oracle_lib as O
    if channels:
        parts = [time_cpu_oracle((wl[0], wl[1], m, c, wl[4], wl[5]), budget_s / len(channels))
                 for (m, c) in channels]
        return {"value": 1.0 / sum(1.0 / p["value"] for p in parts), "unit": "samples/s",
                "cores": parts[0]["cores"], "kind": "port",
                "sample": "one run per channel: " + "; ".join(p["sample"] for p in parts),
                "orbit_points_per_s": None}
    w, h, m, c, canvas, _ = wl
    n, dt = 1 << 22, 0.0
    for _ in range(3):  # grow the sample until it is worth ~budget_s of CPU work
        t0 = time.perf_counter()
        _, cnt, threads = O.render(w, h, m, c, 1337, 0, n, canvas=canvas)
        dt = time.perf_counter() - t0
        if dt >= 0.6 * budget_s or n >= (1 << 31):
            break
        n = int(min(n * max(budget_s / max(dt, 1e-3), 2.0), 1 << 31))
    return {"value": n / dt, "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": "%d samples of the same workload (seed 1337, indices 0..), %.1f s, OpenMP "
                      "oracle/liboracle.so" % (n, dt),
            "orbit_points_per_s": cnt["orbit_points"] / dt}


def run_reference_binary(wl, seconds, device=0):
    """One run of the unmodified reference program for `seconds` (its -t); returns
    (samples/s, passes, seconds it printed) or None if the binary is missing / failed."""
    import oracle_lib as O
    if not os.path.exists(O.REF_BINARY):
        return None
    w, h, m, c, canvas, _ = wl
    with tempfile.TemporaryDirectory() as tmp:
        cmd = [O.REF_BINARY, "-d", str(device), "-w", str(w), "-h", str(h), "-m", str(m), "-c",
               str(c), "--min-real", repr(canvas[0]), "--max-real", repr(canvas[1]),
               "--min-imag", repr(canvas[2]), "--max-imag", repr(canvas[3]), "-t", str(seconds),
               "-g", "-1", "-o", os.path.join(tmp, "ref.pgm")]
        # canvas flags re-validate immediately (cudabrot.cu:726-748): put the max flags first when
        # the new min would not be below the default max
        if canvas[0] >= 2.0 or canvas[2] >= 2.0:
            return None
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=seconds + 600)
        except (OSError, subprocess.TimeoutExpired):
            return None
        mt = re.search(r"(\d+) Buddhabrot passes took ([0-9.]+) seconds", r.stdout)
        if r.returncode != 0 or not mt:
            log("reference binary failed:", r.stdout[-400:], r.stderr[-400:])
            return None
        passes, secs = int(mt.group(1)), float(mt.group(2))
        return passes * REFERENCE_PASS / secs, passes, secs


def make_config(name, wl):
    """The same dict for both arms (the driver compares them)."""
    w, h, m, c, canvas, _ = wl
    return {"workload": "%s: %dx%d canvas %s, max-iter %d, min-cutoff %d" %
            (name, w, h, list(canvas), m, c), "seed": 1337,
            "l2": "cold between timed steps (native arm: 256 MiB flush write; reference arm: one "
                  "process per step)",
            "inputs": "candidate samples are drawn on the device (Philox / the reference's XORWOW)"}


def reference_arm(args, wl, rank):
    """--impl reference: rank 0 alone runs and prints; other ranks exit 0."""
    if rank != 0:
        return 0
    w, h, m, c, canvas, _ = wl
    secs = 5.0
    runs = []
    for i in range(args.warmup + args.steps):
        r = run_reference_binary(wl, secs)
        if r is None:
            runs = None
            break
        if i >= args.warmup:
            runs.append(r)
    config = make_config(args.workload, wl)
    if runs:
        samples = sum(r[1] for r in runs) * REFERENCE_PASS
        t = sum(r[2] for r in runs)
        value = samples / t
        line = {"metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / len(runs),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference", "config": config,
                "reference_kind": "unmodified cudabrot.cu built for sm_100a "
                                  "(oracle/_ref/cudabrot_ref), one B200, -t %.0f per step; time as "
                                  "printed by the program (includes its final D2H copy)" % secs,
                "cpu_baseline": {"value": value, "unit": "samples/s", "cores": 1,
                                 "kind": "reference",
                                 "sample": "%d passes of 13107200 samples in %.1f s on 1 GPU (the "
                                           "reference has no CPU path)" % (sum(r[1] for r in runs), t)},
                "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        # cudabrot has no CPU implementation; for completeness the CPU restatement of its kernel
        # logic (OpenMP, all host threads) on a bounded sample of the same workload
        line["cpu_port"] = time_cpu_oracle(wl, budget_s=8.0)
    else:
        # no GPU build of the reference available: time the CPU port with all host threads
        cb = time_cpu_oracle(wl, budget_s=10.0 * max(1, args.steps))
        line = {"metric": METRIC, "value": cb["value"], "unit": "samples/s", "n_gpus": 1,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": None,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "impl": "reference", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


# ---- roofline model --------------------------------------------------------------------------
# FP64-pipe warp instructions the render kernel EXECUTES, as lane-instructions per unit of work,
# calibrated against ncu (executed DFMA / DMUL / DADD / DSETP of render_persistent_kernel from the
# source pages of `ncu --set full` captures over 2^30-sample launches of the final round-2 kernel,
# profiles/r02_summary.md): a per candidate (the exact tiers' coordinates, cardioid/bulb test,
# re-computed prefixes and tested steps for the 17.5 % of the candidates the FP32 pre-classification
# does not retire), b per escape-pass iteration actually executed (4 per unchecked deep step, 7 per
# tested step, idle lanes and rolled-back rounds included), c per recorded orbit point (step 4 +
# one-rounding binning 2, the exact-binning sliver), d per candidate in the build with the cycle
# certificate.  Fit: config 1 35.0 (model 35.0), config 2 61.7 (61.7), config 3 62.2 (62.2),
# config 4 70.0 (70.0) per candidate.
FP64_MODEL = {"per_candidate": 18.2, "per_executed_iteration": 4.31, "per_orbit_point": 6.2,
              "certificate_per_candidate": 3.2}
# All warp instructions the render kernel executes per unit of work (smsp__inst_executed of the same
# captures: config 1 5.95, config 2 7.06, config 4 9.03 per candidate; the 72-register build of the
# tiled config 3 runs 7 % above the fit: 8.92).  Used for the issue-slot figure: a sub-partition
# issues at most one warp instruction per cycle.
INSTR_MODEL = {"per_candidate": 4.49, "per_executed_iteration": 0.243, "per_orbit_point": 1.52,
               "certificate_per_candidate": 0.11, "tiled_build_factor": 1.07}
HW_FP64_LANES = 148 * 64  # FP64 lanes of one B200: the hardware issue rate is this x the SM clock


def fp64_lane_instr(S, cnt, scale=1.0, cert=False):
    m = FP64_MODEL
    return ((m["per_candidate"] + (m["certificate_per_candidate"] if cert else 0.0)) * S +
            m["per_executed_iteration"] * cnt["executed_iters"] * scale +
            m["per_orbit_point"] * cnt["orbit_points"] * scale)


def make_roofline(S, cnt, scale, t_s, fp64_peak, sm_mhz, workload, cert=False, tiled=False):
    """roofline of the dominant kernel (render_persistent_kernel): executed FP64 lane-instructions
    per second against the in-run DFMA probe; the reference-dataflow figure of SURVEY.md 8(d)
    (14 S + 8 E + 8 P, with E what the reference would have to execute) is kept beside it."""
    lane = fp64_lane_instr(S, cnt, scale, cert)
    lane_ref = 14 * S + 8 * cnt["escape_iters"] * scale + 8 * cnt["orbit_points"] * scale
    hw = HW_FP64_LANES * (sm_mhz or 1965.0) * 1e6
    im = INSTR_MODEL
    warp_instr = ((im["per_candidate"] + (im["certificate_per_candidate"] if cert else 0.0)) * S +
                  im["per_executed_iteration"] * cnt["executed_iters"] * scale +
                  im["per_orbit_point"] * cnt["orbit_points"] * scale)
    if tiled:
        warp_instr *= im["tiled_build_factor"]
    avail_cycles = t_s * 148 * 4 * (sm_mhz or 1965.0) * 1e6
    issue_port = {"frac": warp_instr / avail_cycles,
                  "warp_instructions_per_candidate": warp_instr / S,
                  "cycles_per_candidate_measured": avail_cycles / S,
                  "model": "issue-slot utilisation = warp instructions the kernel executes (ncu-"
                           "calibrated, bench.py INSTR_MODEL) over 148 SM x 4 sub-partitions x SM "
                           "clock x time: a sub-partition issues at most one warp instruction per "
                           "cycle (ncu smsp__issue_active of the same kernel: 71-73 %).  The kernel "
                           "is bound by instruction issue with the half-rate integer pipe (IMAD, "
                           "LOP3, ISETP: Philox is 84 of the ~165 cycles of a sampler batch) close "
                           "behind; the FP64 pipe is 22-35 % busy and its dependency chains largely "
                           "run in slots the other warps leave free (round 2: removing 57 % of the "
                           "executed iterations gained 2 %; profiles/r02_summary.md)"}
    return {
        "bound": "fp64-pipe", "unit": "Tlane-instr/s",
        "achieved": lane / t_s / 1e12, "peak": fp64_peak / 1e12, "frac": lane / t_s / fp64_peak,
        "frac_of_hw_issue_rate": lane / t_s / hw,
        "hw_issue_rate": hw / 1e12,
        "traffic": NCU_DRAM_BYTES_PER_2P30_LAUNCH.get(workload),
        "traffic_note": "DRAM bytes (read+write) of one render_persistent_kernel launch over 2^30 "
                        "samples of this workload (ncu --set full, profiles/); algorithmic memory "
                        "traffic is 4 B per in-canvas increment at L2",
        "peak_source": "in-run independent-DFMA probe (buddha_probe_fp64_peak; reads ~92 % of "
                       "148 SM x 64 lanes x SM clock = hw_issue_rate); MEASURED_PEAKS.json has no "
                       "FP64 figure",
        "numerator": "FP64-pipe lane-instructions the kernel executes: %.1f per candidate + %.2f "
                     "per executed escape iteration + %.1f per orbit point (+ 3.2 per candidate "
                     "with the cycle certificate; ncu-calibrated, "
                     "bench.py FP64_MODEL)" % (FP64_MODEL["per_candidate"],
                                                FP64_MODEL["per_executed_iteration"],
                                                FP64_MODEL["per_orbit_point"]),
        "issue_port": issue_port,
        "frac_reference_dataflow": lane_ref / t_s / fp64_peak,
        "reference_dataflow_note": "14 S + 8 E + 8 P with E = the iterations the REFERENCE must run "
                                   "for the same output (SURVEY.md 8(d)); a speed-up factor of the "
                                   "exact periodicity shortcut and the 4-instruction step, not a "
                                   "utilisation",
    }


def make_red_roofline(r, cnt, scale, t_s, hist_bytes, tiled, workload):
    """Second bound: the histogram reductions.  Direct scatter: the probe's random
    red.global.add.u32 rate over a footprint of the histogram's size.  Tiled scatter (histograms far
    beyond L2): every increment is appended to a list (4 B written, 4 B read back) and applied as a
    reduction inside an L2-resident 64 MB tile, so the ceiling is the in-L2 reduction rate."""
    inc = cnt["increments"] * scale
    foot = min(hist_bytes, 64 << 20) if tiled else hist_bytes
    copies = 1
    if hist_bytes <= (128 << 10):   # the library's rule for privatised copies (buddha_api.cu)
        copies = min(64, (4 << 20) // hist_bytes)
        foot = hist_bytes * copies
    peak = r.probe_red_peak(foot)
    out = {"bound": "l2-red", "unit": "Gred/s", "achieved": inc / t_s / 1e9, "peak": peak / 1e9,
           "frac": inc / t_s / peak, "footprint_bytes": hist_bytes, "tiled_scatter": bool(tiled),
           "peak_source": "in-run probe: red.global.add.u32 to uniformly random cells of a %d MB "
                          "array (%s)" % (foot >> 20, "one L2-resident tile" if tiled else
                                          "the histogram's size"),
           "algorithmic_bytes_per_increment": 4}
    if copies > 1:
        out["privatised_copies"] = copies
    if tiled:
        # HBM bytes the tiled pipeline has to move: the list entry written and read back (8 B per
        # increment) plus one read and one write-back of every 64 MB tile per pipeline launch
        # (a launch = 1 render + 1 drain + n_tiles apply kernels)
        n_tiles = -(-hist_bytes // (64 << 20))
        launches = max(cnt["kernel_launches"] * scale / (2 + n_tiles), 1.0)
        model = 8.0 * inc + 2.0 * hist_bytes * launches
        out["hbm_model_bytes_per_increment"] = model / max(inc, 1.0)
        out["hbm_model_gbs"] = model / t_s / 1e9
        out["hbm_model"] = "4 B list write + 4 B list read per increment, plus one read and one " \
                           "write-back of each 64 MB tile per pipeline launch; 4 B algorithmic"
        dram = NCU_DRAM_BYTES_PER_2P30_LAUNCH.get(workload)
        if dram:
            out["ncu_dram_bytes_per_2p30_launch"] = dram
    return out


def e2e_pipeline(r, ranges, host_in, host_hist, host_img, n_ch, shape, rank, world, hist_t, dist):
    """The job through the public C ABI with HOST buffers, copies overlapped with rendering
    (buddha.h "Overlapped host transfers").  Per step: the saved in-progress counts go in
    (buddha_add_histogram_async, H2D from pinned memory, root only), every rank renders its range,
    N > 1: one reduce(sum) to the root and the others start from zero again, the root freezes the
    result (buddha_snapshot) and reads histogram + 16-bit image of step k-1 back (D2H) while step k
    renders.  Returns wall seconds (caller brackets with barriers)."""
    import torch
    h, w = shape
    t0 = time.perf_counter()
    pending = False
    for (f, n) in ranges:
        if rank == 0:
            r.add_histogram_async(host_in)
        r.render_samples_async(f, n)
        if rank == 0 and pending:
            r.read_snapshot(host_hist)
            for ch in range(n_ch):
                r.tonemap_snapshot(1.0, True, out=host_img.reshape(n_ch, h, w)[ch], channel=ch)
        r.sync()
        if world > 1:
            dist.reduce(hist_t, dst=0, op=dist.ReduceOp.SUM)
            if hist_t.is_cuda:          # (the CPU test drives this loop with gloo tensors)
                torch.cuda.synchronize()
            if rank != 0:
                r.clear()
        if rank == 0:
            r.snapshot()
            pending = True
    if rank == 0 and pending:
        r.read_snapshot(host_hist)
        for ch in range(n_ch):
            r.tonemap_snapshot(1.0, True, out=host_img.reshape(n_ch, h, w)[ch], channel=ch)
    r.sync()
    return time.perf_counter() - t0


def pinned(n, dtype):
    import numpy as np
    import torch
    tdt = {"uint32": torch.int32, "uint16": torch.int16}[dtype]
    return torch.zeros(n, dtype=tdt).pin_memory().numpy().view(getattr(np, dtype))


class Ctx:
    """What every part of the native arm needs."""
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        self.torch, self.dist = torch, dist
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
        # host-side barrier (gloo): while rank 0 runs another PROCESS on all GPUs (cli_merge_check)
        # the other ranks must not sit in an NCCL barrier kernel -- the GPU would time-slice the
        # two processes and the other process's collectives crawl
        self.host_group = dist.new_group(backend="gloo") if world > 1 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def host_barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.host_group)

    def allmax(self, *vals):
        if self.world == 1:
            return list(vals)
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def allsum_counters(self, cnt, extra=()):
        keys = ["rejected", "hit_max", "too_early", "accepted", "escape_iters", "orbit_points",
                "increments", "executed_iters", "shortcut_hits", "kernel_launches", "exact_bins"]
        vals = [cnt[k] for k in keys] + list(extra)
        if self.world > 1:
            t = self.torch.tensor(vals, dtype=self.torch.int64, device="cuda")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
            vals = t.tolist()
        out = dict(cnt)
        out.update(zip(keys, vals[:len(keys)]))
        return out, vals[len(keys):]

    def bcast_int(self, v):
        if self.world == 1:
            return int(v)
        t = self.torch.tensor([int(v)], dtype=self.torch.int64, device="cuda")
        self.dist.broadcast(t, src=0)
        return int(t[0])


def run_workload(cx, name, steps, warmup, per_gpu, min_seconds=0.0, e2e_steps=None,
                 baselines=False):
    """Weak-scaling run of one BASELINE workload: every (step, rank) renders a fresh block of
    `per_gpu` sample indices; device-timed per step with an L2 flush in between; then the e2e
    pipeline.  steps = 0: as many steps as `min_seconds` of device time need.  Returns the result
    dict on rank 0 (None elsewhere)."""
    import numpy as np
    import cudabrot_b200 as B
    from cudabrot_b200.sharding import step_range
    torch, dist = cx.torch, cx.dist
    rank, world = cx.rank, cx.world
    w, h, m, c, canvas, _ = WORKLOADS[name]
    channels = CHANNELS.get(name)
    flags = B.F_NO_SHORTCUT if cx.args.no_shortcut else 0
    r = B.Renderer(w, h, m, c, canvas=canvas, seed=1337, device=cx.local_rank, flags=flags,
                   channels=channels)
    n_ch = len(channels) if channels else 1
    cells = w * h * n_ch
    hist_t = r.histogram_as_tensor()
    hist_bytes = hist_t.numel() * 4
    tiled = hist_bytes >= (640 << 20)

    # ---- warm-up (untimed): also calibrates the tile lists of a tiled context ----------------
    for k in range(max(warmup, 1)):
        f, n = step_range(k, rank, world, min(per_gpu, 1 << 32) if steps == 0 else per_gpu,
                          first=1 << 56)
        r.render_samples(f, n)
        t_first = r.last_render_ms() * per_gpu / n
    if steps == 0:
        steps = cx.bcast_int(max(2, int(np.ceil(min_seconds * 1e3 / max(t_first, 1e-3)))))
    if world > 1:
        # warm-up of the one collective too: NCCL sets up its buffers for a new message size at the
        # first call (a 1.2 GB reduce once took 143 ms instead of 17)
        dist.reduce(hist_t, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    r.clear()
    r.reset_counters()
    fp64_peak = r.probe_fp64_peak()

    sampler = ClockSampler(cx.local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()  # nvidia-smi needs a moment before its first line
    # ---- timed region: `steps` steps, device-timed per step --------------------------------
    cx.barrier()
    sampler.mark_start()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    for k in range(steps):
        cx.flush.zero_()                    # L2 flush between timed steps (not timed)
        torch.cuda.synchronize()
        f, n = step_range(k, rank, world, per_gpu)
        r.render_samples(f, n)
        dev_ms += r.last_render_ms()        # CUDA events on libbuddha's own stream
    merge_ms = 0.0
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        dist.reduce(hist_t, dst=0, op=dist.ReduceOp.SUM)   # the single collective: ncclReduce(sum)
        e1.record()
        torch.cuda.synchronize()
        merge_ms = e0.elapsed_time(e1)
    cx.barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    clocks = sampler.stop() if rank == 0 else None
    cnt_local = r.counters()
    inc_local = sum(r.channel_counters(k)["increments"] for k in range(n_ch)) if channels \
        else cnt_local["increments"]
    total_ms, wall_ms, merge_max = cx.allmax(dev_ms + merge_ms, wall_ms, merge_ms)
    cnt, (inc_total,) = cx.allsum_counters(cnt_local, extra=(inc_local,))
    samples_total = per_gpu * steps * world
    value = samples_total / (total_ms * 1e-3)
    if rank == 0:
        # sanity: the merged histogram holds exactly the increments all ranks counted (fused
        # contexts keep one device histogram per band: one cell increment per in-canvas point)
        merged_sum = int(hist_t.view(torch.int32).to(torch.int64).bitwise_and(0xFFFFFFFF).sum())
        if merged_sum != cnt["increments"]:
            raise RuntimeError("%s: histogram sum %d != increments %d" %
                               (name, merged_sum, cnt["increments"]))

    # ---- e2e: the same kind of steps through the C ABI with host buffers, copies overlapped ---
    n_e2e = min(steps, e2e_steps or steps)
    r.clear()
    host_in = host_hist = host_img = None
    if rank == 0:
        host_in = pinned(cells, "uint32")
        host_hist = pinned(cells, "uint32").reshape((n_ch, h, w) if channels else (h, w))
        host_img = pinned(cells, "uint16")
    ranges = [step_range(k, rank, world, per_gpu, first=1 << 57) for k in range(n_e2e)]
    inc_before = r.counters()["increments"]
    # one short untimed pass first: the staging / snapshot / scratch buffers are allocated on first use
    e2e_pipeline(r, [step_range(0, rank, world, 1 << 22, first=1 << 58)], host_in, host_hist, host_img,
                 n_ch, (h, w), rank, world, hist_t, dist)
    cx.barrier()
    e2e_s = e2e_pipeline(r, ranges, host_in, host_hist, host_img, n_ch, (h, w), rank, world,
                         hist_t, dist)
    cx.barrier()
    (e2e_s,) = cx.allmax(e2e_s)
    e2e_value = per_gpu * n_e2e * world / e2e_s
    # the histogram that came back over PCIe must hold exactly what all ranks counted meanwhile
    _, (inc_e2e,) = cx.allsum_counters(r.counters(), extra=(r.counters()["increments"] - inc_before,))
    if rank == 0 and not channels:
        got = int(host_hist.sum(dtype=np.uint64))
        if got != inc_e2e:
            raise RuntimeError("%s: e2e histogram sum %d != increments %d" % (name, got, inc_e2e))

    if rank != 0:
        r.close()
        return None
    scale = 1.0 / world  # counters were summed over ranks; the rooflines are per GPU
    t_s = dev_ms * 1e-3
    S0 = per_gpu * steps
    res = {
        "workload": name, "value": value, "samples_per_s": value, "unit": "samples/s",
        "steps": steps, "samples_per_step_per_gpu": per_gpu, "ms_per_step": total_ms / steps,
        "orbit_points_per_s": cnt["orbit_points"] / (total_ms * 1e-3),
        "increments_per_s": inc_total / (total_ms * 1e-3),
        "wall_ms_per_step": wall_ms / steps, "merge_ms": merge_max,
        "counters": {k: cnt[k] for k in ("rejected", "hit_max", "too_early", "accepted",
                                         "escape_iters", "executed_iters", "orbit_points",
                                         "increments", "shortcut_hits", "exact_bins")},
        "gpu_launches": cnt["kernel_launches"] // world,
        "e2e": {"value": e2e_value, "unit": "samples/s", "steps": n_e2e,
                "h2d_bytes_per_step": cells * 4, "d2h_bytes_per_step": cells * 4 + cells * 2,
                "frac_of_device_rate": e2e_value / value,
                "what": "per step through the C ABI, pinned host buffers, wall clock: "
                        "add_histogram_async (H2D) + render_samples_async%s + snapshot + "
                        "read_snapshot (D2H) + tonemap_snapshot_u16 (D2H); the copies of step k-1 "
                        "overlap the render of step k" %
                        (" + ncclReduce to rank 0 (only rank 0 copies)" if world > 1 else "")},
        "clocks": clocks,
        "roofline": make_roofline(S0, cnt, scale, t_s, fp64_peak,
                                  clocks["sm_mhz"] if clocks else None, name,
                                  cert=(not channels and not tiled and m >= 10000 and
                                        not cx.args.no_shortcut),
                                  tiled=tiled),
    }
    res["roofline_red"] = make_red_roofline(r, cnt, scale, t_s, hist_bytes, tiled, name)
    if channels:
        res["channels"] = channels
    r.close()
    if baselines and world == 1:
        wl = WORKLOADS[name]
        refs = [run_reference_binary((w, h, mk, ck, canvas, 0), 3.0)
                for (mk, ck) in (channels or [(m, c)])]
        if all(refs):
            res["reference_cuda"] = {"value": 1.0 / sum(1.0 / x[0] for x in refs),
                                     "unit": "samples/s", "passes": [x[1] for x in refs],
                                     "seconds": [x[2] for x in refs],
                                     "what": "unmodified cudabrot.cu built for sm_100a, same "
                                             "workload, this GPU, -t 3 (one run per channel)"}
            res["vs_reference_cuda"] = res["value"] / res["reference_cuda"]["value"]
    return res


def run_strong(cx, name="cfg3_m20000", n_total=1 << 39):
    """Strong scaling on the render north_star names for 8 GPUs: a FIXED range of n_total sample
    indices split contiguously over the ranks, one ncclReduce(sum) of the 1.6 GB histograms timed
    inside the job, and the digest of the merged histogram -- which must be the same for every
    number of GPUs (histogram(N GPUs) == histogram(1 GPU))."""
    import cudabrot_b200 as B
    from cudabrot_b200.sharding import contiguous_split
    torch, dist = cx.torch, cx.dist
    rank, world = cx.rank, cx.world
    w, h, m, c, canvas, _ = WORKLOADS[name]
    r = B.Renderer(w, h, m, c, canvas=canvas, seed=1337, device=cx.local_rank)
    hist_t = r.histogram_as_tensor()
    r.render_samples(1 << 56, 1 << 26)      # warm-up: sizes the tile lists, loads the kernels
    if world > 1:
        dist.reduce(hist_t, dst=0, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
    r.clear()
    r.reset_counters()
    sampler = ClockSampler(cx.local_rank)
    if rank == 0:
        sampler.start()
        sampler.wait_first()
    f, n = contiguous_split(0, n_total, rank, world)
    cx.barrier()
    sampler.mark_start()
    wall0 = time.perf_counter()
    r.render_samples(f, n)
    render_ms = r.last_render_ms()
    merge_ms = 0.0
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.reduce(hist_t, dst=0, op=dist.ReduceOp.SUM)
        e1.record()
        torch.cuda.synchronize()
        merge_ms = e0.elapsed_time(e1)       # includes waiting for the slowest rank's render
    cx.barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    clocks = sampler.stop() if rank == 0 else None
    total_ms, render_max, wall_ms = cx.allmax(render_ms + merge_ms, render_ms, wall_ms)
    cnt, _ = cx.allsum_counters(r.counters())
    res = None
    if rank == 0:
        res = {"workload": name, "n_total": n_total, "split": "contiguous, rank r of N",
               "ms": total_ms, "render_ms_max": render_max, "merge_ms": merge_ms,
               "wall_ms": wall_ms, "samples_per_s": n_total / (total_ms * 1e-3),
               "orbit_points_per_s": cnt["orbit_points"] / (total_ms * 1e-3),
               "hist_fnv": "%016x" % r.digest(0),
               "hist_fnv_what": "buddha_histogram_digest of the merged histogram on rank 0 "
                                "(blocked FNV-1a-64, include/buddha.h)",
               "merge_bytes": hist_t.numel() * 4,
               "counters": {k: cnt[k] for k in ("accepted", "orbit_points", "increments",
                                                "executed_iters")},
               "clocks": clocks, "timing": "render: CUDA events on the library's stream; merge: "
               "CUDA events around ncclReduce; ms = max over ranks of their sum"}
    r.close()
    return res


def cli_merge_check(cx):
    """buddha_merge (the in-process ncclReduce behind `cudabrot --gpus N`): the PGM written with
    --gpus N must equal the one written with --gpus 1 for the same sample range."""
    import hashlib
    from cudabrot_b200 import capi
    if cx.rank != 0:
        return None
    out = {"gpus": cx.world}
    with tempfile.TemporaryDirectory() as tmp:
        shas = []
        for g in sorted({cx.world, 1}, reverse=True):
            pgm = os.path.join(tmp, "g%d.pgm" % g)
            cmd = [capi.CLI_PATH, "--gpus", str(g), "-w", "2000", "-h", "2000", "-m", "2000", "-c",
                   "20", "--samples", str(1 << 30), "-o", pgm]
            try:
                p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            except (OSError, subprocess.TimeoutExpired) as e:
                out["error"] = str(e)
                return out
            if p.returncode != 0 or not os.path.exists(pgm):
                out["error"] = (p.stdout + p.stderr)[-300:]
                return out
            shas.append(hashlib.sha256(open(pgm, "rb").read()).hexdigest()[:16])
        out["pgm_sha_gpusN"], out["pgm_sha_gpus1"] = shas[0], shas[-1]
        out["equal"] = shas[0] == shas[-1]
        out["what"] = "bin/cudabrot --gpus N vs --gpus 1, 2000x2000 -m 2000 -c 20 --samples 2^30"
    return out


def native_arm(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        tmp = torch.zeros(1024, dtype=torch.int32, device="cuda")
        dist.reduce(tmp, dst=0, op=dist.ReduceOp.SUM)   # initialises NCCL outside any timed region
        torch.cuda.synchronize()
    cx = Ctx(args, rank, world, local_rank)
    w, h, m, c, canvas, default_step = wl
    per_gpu = args.samples_per_step or default_step

    head = run_workload(cx, args.workload, args.steps, args.warmup, per_gpu)
    extras, strong, cli = [], None, None
    if not args.no_extras:
        strong = run_strong(cx, n_total=args.strong_samples)
        cx.host_barrier()
        cli = cli_merge_check(cx)
        cx.host_barrier()
        for name in ("cfg1", "cfg3", "cfg4", "cfg5"):
            if name == args.workload:
                continue
            res = run_workload(cx, name, 0, 1, WORKLOADS[name][5], min_seconds=1.6, e2e_steps=6,
                               baselines=not args.skip_baselines)
            if res:
                extras.append(res)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return 0

    line = {
        "metric": METRIC, "value": head["value"], "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": head["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": make_config(args.workload, wl),
        "run": {"samples_per_step_per_gpu": per_gpu,
                "parallelism": "%d x disjoint Philox ranges%s" %
                (world, " + 1 ncclReduce(sum) at the end (timed)" if world > 1 else ""),
                "shortcut": not args.no_shortcut},
        "orbit_points_per_s": head["orbit_points_per_s"],
        "increments_per_s": head["increments_per_s"],
        "wall_ms_per_step": head["wall_ms_per_step"], "merge_ms": head["merge_ms"],
        "counters": head["counters"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
        "clocks": head["clocks"], "roofline": head["roofline"], "roofline_red": head["roofline_red"],
    }
    line["e2e"]["merge_ms"] = head["merge_ms"]
    # the driver keeps `roofline`, `e2e`, `config`, `clocks`, `cpu_baseline` whole: the other
    # BASELINE workloads and the strong-scaling job ride inside `roofline` as well as at top level
    if strong:
        line["strong_cfg3_m20000"] = strong
        line["roofline"]["strong_cfg3_m20000"] = strong
    if cli:
        line["cli_merge"] = cli
        line["roofline"]["cli_merge"] = cli
    if extras:
        line["workloads"] = extras
        line["roofline"]["workloads"] = extras
    if world == 1 and not args.skip_baselines:
        channels = CHANNELS.get(args.workload)
        line["cpu_baseline"] = time_cpu_oracle(wl, channels=channels)
        refs = [run_reference_binary((w, h, mk, ck, canvas, 0), 5.0)
                for (mk, ck) in (channels or [(m, c)])]
        if all(refs):
            line["reference_cuda"] = {"value": 1.0 / sum(1.0 / x[0] for x in refs),
                                      "unit": "samples/s", "passes": [x[1] for x in refs],
                                      "seconds": [x[2] for x in refs],
                                      "what": "unmodified cudabrot.cu built for sm_100a, same "
                                              "workload, this GPU, -t 5 (one run per channel)"}
    emit(line)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--samples-per-step", type=int, default=0)
    ap.add_argument("--no-shortcut", action="store_true")
    ap.add_argument("--skip-baselines", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="only the headline workload: no strong-scaling job, no other workloads")
    ap.add_argument("--strong-samples", type=int, default=1 << 39)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if not (args.impl != "reference" and world != args.gpus and world == 1 and args.gpus > 1):
        claim_stdout()  # (not in the parent that only re-launches itself under torchrun)
    if args.impl == "reference":
        return reference_arm(args, wl, rank)
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-launch ourselves one process per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return native_arm(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
