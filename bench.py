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
    "cfg3": (20000, 20000, 2000, 20, FULL, 1 << 32),
    "cfg3_m20000": (20000, 20000, 20000, 20, FULL, 1 << 32),
    "cfg4": (8000, 4000, 5000, 20, (0.0, 1.0, 0.0, 0.5), 1 << 32),
    "cfg5a": (10000, 10000, 100, 20, FULL, 1 << 32),
    "cfg5b": (10000, 10000, 1000, 20, FULL, 1 << 32),
    "cfg5c": (10000, 10000, 20000, 20, FULL, 1 << 32),
    # config 5 as ONE fused pass: every candidate is rendered once into the three channels
    "cfg5": (10000, 10000, 20000, 20, FULL, 1 << 32),
    # the same trio on a canvas whose three histograms fit L2 together (192 MB), fused and apart
    "cfg5_4k": (4000, 4000, 20000, 20, FULL, 1 << 32),
    "cfg5_4k_a": (4000, 4000, 100, 20, FULL, 1 << 32),
    "cfg5_4k_b": (4000, 4000, 1000, 20, FULL, 1 << 32),
    "cfg5_4k_c": (4000, 4000, 20000, 20, FULL, 1 << 32),
}
# fused multi-channel workloads: [(max-iter, min-cutoff)] per channel (BASELINE.json configs[4])
CHANNELS = {"cfg5": [(100, 20), (1000, 20), (20000, 20)],
            "cfg5_4k": [(100, 20), (1000, 20), (20000, 20)]}
REFERENCE_PASS = 13107200  # 512 blocks * 512 threads * 50 samples, cudabrot.cu:20,23,34
# dram__bytes_read.sum + dram__bytes_write.sum of ONE render_persistent_kernel launch over 2^30
# samples, from the `ncu --set full` captures summarised in profiles/r01_summary.md
NCU_DRAM_BYTES_PER_2P30_LAUNCH = {"cfg1": 4037888 + 1280, "cfg2": 37888 + 256,
                                  "cfg3": 208894976 + 5066775000}
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
    config = {"workload": "%s: %dx%d canvas %s, max-iter %d, min-cutoff %d" %
              (args.workload, w, h, list(canvas), m, c), "seed": 1337}
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


def native_arm(args, wl, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    import cudabrot_b200 as B
    from cudabrot_b200.sharding import merge_to_root, step_range

    w, h, m, c, canvas, default_step = wl
    per_gpu = args.samples_per_step or default_step
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    flags = B.F_NO_SHORTCUT if args.no_shortcut else 0
    channels = CHANNELS.get(args.workload)
    r = B.Renderer(w, h, m, c, canvas=canvas, seed=1337, device=local_rank, flags=flags,
                   channels=channels)
    n_ch = len(channels) if channels else 1
    cells = w * h * n_ch
    hist_t = r.histogram_as_tensor()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- warm-up (untimed), incl. one collective so NCCL is initialised -------------------
    for k in range(args.warmup):
        f, n = step_range(k, rank, world, per_gpu, first=1 << 56)
        r.render_samples(f, n)
    if world > 1:
        tmp = torch.zeros(1024, dtype=torch.int32, device="cuda")
        merge_to_root(tmp)
    r.clear()
    r.reset_counters()
    fp64_peak = r.probe_fp64_peak()

    # ---- timed region: K steps (+ the one merge), device-timed per step ---------------------
    barrier()
    sampler.mark_start()
    wall0 = time.perf_counter()
    dev_ms = 0.0
    for k in range(args.steps):
        flush.zero_()                       # L2 flush between timed steps (not timed)
        torch.cuda.synchronize()
        f, n = step_range(k, rank, world, per_gpu)
        r.render_samples(f, n)
        dev_ms += r.last_render_ms()        # CUDA events on libbuddha's own stream
    merge_ms = 0.0
    if world > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        merge_to_root(hist_t)               # the single collective: ncclReduce(sum) to rank 0
        e1.record()
        torch.cuda.synchronize()
        merge_ms = e0.elapsed_time(e1)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - wall0)
    clocks = sampler.stop() if rank == 0 else None
    cnt = r.counters()
    # cells incremented: fused contexts add a point to every accepting channel
    inc_local = sum(r.channel_counters(k)["increments"] for k in range(n_ch)) if channels \
        else cnt["increments"]
    increments_total = inc_local

    total_ms = dev_ms + merge_ms
    if world > 1:
        t = torch.tensor([total_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, wall_ms = float(t[0]), float(t[1])
        keys = ["rejected", "hit_max", "too_early", "accepted", "escape_iters", "orbit_points",
                "increments", "executed_iters", "shortcut_hits", "kernel_launches", "exact_bins"]
        ct = torch.tensor([cnt[k] for k in keys], dtype=torch.int64, device="cuda")
        dist.all_reduce(ct, op=dist.ReduceOp.SUM)
        for k, v in zip(keys, ct.tolist()):
            cnt[k] = v
        it = torch.tensor([inc_local], dtype=torch.int64, device="cuda")
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        increments_total = int(it[0])
    samples_total = per_gpu * args.steps * world
    value = samples_total / (total_ms * 1e-3)

    # sanity: the merged histogram holds exactly the increments all ranks counted
    if rank == 0:
        merged_sum = int(hist_t.view(torch.int32).to(torch.int64).bitwise_and(0xFFFFFFFF).sum())
        # (fused contexts keep one device histogram per band: every in-canvas point is one cell
        # increment there, cnt["increments"]; the per-channel sum increments_total counts a point
        # once per accepting channel)
        if merged_sum != cnt["increments"]:
            raise RuntimeError("histogram sum %d != increments %d" % (merged_sum, cnt["increments"]))

    # ---- e2e: the same steps through the C ABI with HOST buffers ---------------------------
    # per step: H2D of the in-progress histogram (the -s buffer, cudabrot.cu:256), render, D2H of
    # the histogram (:496) and the tone-mapped 16-bit image (:500); pinned host memory
    host_hist = torch.zeros(cells, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
    host_img = torch.zeros(cells, dtype=torch.int16).pin_memory().numpy().view(np.uint16)
    barrier()
    e2e_t0 = time.perf_counter()
    for k in range(args.steps):
        f, n = step_range(k, rank, world, per_gpu, first=1 << 57)
        r.load_histogram(host_hist)
        r.render_samples(f, n)
        r.read_histogram(host_hist.reshape((n_ch, h, w) if channels else (h, w)))
        for ch in range(n_ch):
            r.tonemap(1.0, True, out=host_img.reshape(n_ch, h, w)[ch], channel=ch)
    barrier()
    e2e_s = time.perf_counter() - e2e_t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    e2e_value = samples_total / e2e_s

    if rank != 0:
        r.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline: FP64-pipe issue rate; peak = in-run DFMA probe ------------
    S = samples_total
    t_s = dev_ms * 1e-3  # render kernels only (max over ranks not needed: rank 0's own share)
    S0 = per_gpu * args.steps
    scale = 1.0 / world  # counters were summed over ranks; roofline is per GPU
    lane_exec = (14 * S0 + 8 * cnt["executed_iters"] * scale + 8 * cnt["orbit_points"] * scale)
    lane_ref = (14 * S0 + 8 * cnt["escape_iters"] * scale + 8 * cnt["orbit_points"] * scale)
    roofline = {
        "bound": "fp64-pipe", "unit": "Tlane-instr/s",
        "achieved": lane_exec / t_s / 1e12, "peak": fp64_peak / 1e12,
        "frac": lane_exec / t_s / fp64_peak,
        "traffic": NCU_DRAM_BYTES_PER_2P30_LAUNCH.get(args.workload),
        "traffic_note": "DRAM bytes (read+write) of one render_persistent_kernel launch over 2^30 "
                        "samples of this workload, ncu --set full (profiles/r01_summary.md); the "
                        "kernel is FP64/issue bound, its algorithmic memory traffic is 4 B per "
                        "in-canvas increment at L2",
        "peak_source": "in-run independent-DFMA probe (buddha_probe_fp64_peak); MEASURED_PEAKS.json "
                       "has no FP64 figure",
        "numerator": "reference-dataflow FP64 instructions (14*S + 8*E + 8*P, SURVEY.md 8(d)) with "
                     "E = iterations actually executed",
        "achieved_reference_work": lane_ref / t_s / 1e12,
        "frac_reference_work": lane_ref / t_s / fp64_peak,
        "note": "reference_work counts the iterations the reference must run for the same output "
                "(E includes max-iter for never-escaping samples); the exact periodicity shortcut "
                "and the 4-instruction scaled step make it exceed 1",
    }

    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "%s: %dx%d canvas %s, max-iter %d, min-cutoff %d" %
                   (args.workload, w, h, list(canvas), m, c),
                   "samples_per_step_per_gpu": per_gpu, "seed": 1337,
                   "l2": "256 MiB flush write between timed steps",
                   "parallelism": "%d x disjoint Philox ranges%s" %
                   (world, " + 1 ncclReduce(sum) at the end (timed)" if world > 1 else ""),
                   "shortcut": not args.no_shortcut,
                   **({"channels": channels, "fused": "one pass feeds all channels"}
                      if channels else {})},
        "orbit_points_per_s": cnt["orbit_points"] / (total_ms * 1e-3),
        "increments_per_s": increments_total / (total_ms * 1e-3),
        "wall_ms_per_step": wall_ms / args.steps, "merge_ms": merge_ms,
        "counters": {k: cnt[k] for k in ("rejected", "hit_max", "too_early", "accepted",
                                         "escape_iters", "executed_iters", "orbit_points",
                                         "increments", "shortcut_hits", "exact_bins")},
        "e2e": {"value": e2e_value, "unit": "samples/s",
                "h2d_bytes_per_step": cells * 4, "d2h_bytes_per_step": cells * 4 + cells * 2,
                "what": "per step: load_histogram (H2D) + render_samples + read_histogram (D2H) + "
                        "tonemap_u16 (D2H), pinned host buffers, wall clock"},
        "gpu_launches": cnt["kernel_launches"] // world if world > 1 else cnt["kernel_launches"],
        "clocks": clocks, "roofline": roofline,
    }

    # second roofline: red.global.add.u32 rate vs a probe scattering over the same footprint
    red_peak = r.probe_red_peak(hist_t.numel() * 4)
    line["roofline_red"] = {
        "bound": "l2-red", "unit": "Gred/s", "achieved": cnt["increments"] * scale / t_s / 1e9,
        "peak": red_peak / 1e9, "frac": cnt["increments"] * scale / t_s / red_peak,
        "footprint_bytes": hist_t.numel() * 4,
        "peak_source": "in-run probe: red.global.add.u32 to uniformly random cells of an array of "
                       "the histogram's size (buddha_probe_red_peak)",
        "algorithmic_bytes_per_increment": 4}
    if world == 1 and not args.skip_baselines:
        r.close()
        line["cpu_baseline"] = time_cpu_oracle(wl, channels=channels)
        refs = [run_reference_binary((w, h, mk, ck, canvas, 0), 5.0)
                for (mk, ck) in (channels or [(m, c)])]
        if all(refs):
            line["reference_cuda"] = {"value": 1.0 / sum(1.0 / x[0] for x in refs),
                                      "unit": "samples/s", "passes": [x[1] for x in refs],
                                      "seconds": [x[2] for x in refs],
                                      "what": "unmodified cudabrot.cu built for sm_100a, same "
                                              "workload, this GPU, -t 5 (one run per channel)"}
    else:
        r.close()
    emit(line)
    if world > 1:
        dist.destroy_process_group()
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
