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
    fitted = {"cfg1": (3.11, 0.431, 34.6), "cfg2": (18.03, 0.102, 98.4), "cfg3": (8.81, 1.16, 67.4),
              "cfg4": (12.13, 1.37, 82.3)}
    for name, (e, p, lane) in fitted.items():
        got = bench.fp64_lane_instr(1.0, {"executed_iters": e, "orbit_points": p})
        assert abs(got - lane) / lane < 0.03, (name, got, lane)
