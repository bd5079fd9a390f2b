"""CPU tests of the multi-GPU host logic: disjoint sample ranges + one reduce(sum), exercised with
world_size 2 over gloo; the per-rank "GPU render" is played by the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_contiguous_split_covers_range():
    from cudabrot_b200.sharding import contiguous_split, step_range
    for count in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            pos = 17
            for r in range(world):
                f, n = contiguous_split(17, count, r, world)
                assert f == pos
                pos += n
            assert pos == 17 + count
    seen = set()
    for s in range(3):
        for r in range(4):
            f, n = step_range(s, r, 4, 100, first=5)
            assert n == 100 and f not in seen
            seen.add(f)
    assert sorted(seen) == [5 + 100 * k for k in range(12)]
    with pytest.raises(ValueError):
        contiguous_split(0, 10, 2, 2)


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    from cudabrot_b200.sharding import contiguous_split, merge_to_root
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, n = contiguous_split(1000, 200001, rank, world)
    # preload near-overflow counts on rank 1 to check uint32 wrap-around survives the int32 sum
    base = np.zeros((90, 120), dtype=np.uint32)
    if rank == 1:
        base[0, 0] = 0xFFFFFFFF
        base[0, 1] = 0x7FFFFFFF
    if rank == 0:
        base[0, 0] = 5
        base[0, 1] = 3
    hist, _, _ = O.render(120, 90, 150, 10, 1337, first, n, hist=base, threads=2)
    t = torch.from_numpy(hist.view(np.int32).reshape(-1))
    merge_to_root(t, root=0)
    if rank == 0:
        np.save(out_path, t.numpy().view(np.uint32).reshape(90, 120))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_merge_equals_single_render(oracle, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "merged.npy")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    merged = np.load(out)
    single, _, _ = oracle.render(120, 90, 150, 10, 1337, 1000, 200001)
    expect = single.copy()
    expect[0, 0] += np.uint32(4)            # 5 + 0xFFFFFFFF wraps to 4
    expect[0, 1] += np.uint32(0x80000002)   # 3 + 0x7FFFFFFF
    assert np.array_equal(merged, expect)


def test_bench_arms_share_config_and_model_is_sane():
    """Both bench arms must print the same `config` dict (the driver compares them), and the
    executed-FP64 model must reproduce the ncu counts it was fitted to (profiles/r02_summary.md)."""
    import bench
    for name, wl in bench.WORKLOADS.items():
        assert bench.make_config(name, wl) == bench.make_config(name, wl)
        assert set(bench.make_config(name, wl)) == {"workload", "seed", "l2", "inputs"}
    # (executed iterations and orbit points per candidate, certificate build?, FP64 lane-instructions
    #  per candidate on the source page of the final round-2 captures)
    fitted = {"cfg1": (3.29, 0.431, False, 35.0), "cfg2": (9.16, 0.150, True, 61.7),
              "cfg3": (8.55, 1.166, False, 62.2), "cfg4": (10.05, 1.378, False, 70.0)}
    for name, (e, p, cert, lane) in fitted.items():
        got = bench.fp64_lane_instr(1.0, {"executed_iters": e, "orbit_points": p}, cert=cert)
        assert abs(got - lane) / lane < 0.03, (name, got, lane)


class _OracleRenderer:
    """CPU stand-in with the Renderer calls bench.e2e_pipeline makes, backed by the oracle: lets the
    multi-rank control flow of the e2e loop (root adds saved counts, every rank renders, per-step
    reduce to root, the others start from zero again, snapshot read back one step late) run under
    gloo without a GPU."""
    def __init__(self, oracle, w, h, m, c):
        self.o, self.w, self.h, self.m, self.c = oracle, w, h, m, c
        self.hist = np.zeros((h, w), dtype=np.uint32)
        self.snap = None

    def add_histogram_async(self, saved):
        self.hist += saved.reshape(self.h, self.w)

    def render_samples_async(self, first, n):
        self.o.render(self.w, self.h, self.m, self.c, 1337, first, n, hist=self.hist, threads=2)

    def sync(self):
        pass

    def clear(self):
        self.hist[:] = 0

    def snapshot(self):
        self.snap = self.hist.copy()

    def read_snapshot(self, out):
        out[...] = self.snap

    def tonemap_snapshot(self, gamma, big_endian, out, channel=0):
        img, mx, scale = self.o.tonemap(self.snap, gamma, big_endian=big_endian)
        out[...] = img
        return out, mx, scale


def _e2e_worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import oracle_lib as O
    import bench
    from cudabrot_b200.sharding import step_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    w, h = 96, 64
    r = _OracleRenderer(O, w, h, 200, 10)
    hist_t = torch.from_numpy(r.hist.view(np.int32).reshape(-1))   # aliases the histogram
    saved = np.full(w * h, 3, dtype=np.uint32)
    host_hist = np.zeros((h, w), dtype=np.uint32)
    host_img = np.zeros(w * h, dtype=np.uint16)
    ranges = [step_range(k, rank, world, 20000, first=7) for k in range(3)]
    bench.e2e_pipeline(r, ranges, saved if rank == 0 else None, host_hist, host_img, 1, (h, w),
                       rank, world, hist_t, dist)
    if rank == 0:
        np.save(out_path, host_hist)
        np.save(out_path + ".img.npy", host_img)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_e2e_pipeline_control_flow(oracle, tmp_path):
    """bench.e2e_pipeline at world size 2 (gloo): what rank 0 reads back after the last step is the
    oracle's histogram over all six (step, rank) ranges plus the saved counts of three steps, and
    its tone-mapped image."""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "e2e.npy")
    mp.spawn(_e2e_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    expect = np.full((64, 96), 9, dtype=np.uint32)       # 3 steps x saved counts of 3
    oracle.render(96, 64, 200, 10, 1337, 7, 6 * 20000, hist=expect)   # the six ranges are contiguous
    assert np.array_equal(got, expect)
    img, _, _ = oracle.tonemap(expect, 1.0, big_endian=True)
    assert np.array_equal(np.load(out + ".img.npy").reshape(64, 96), img)
