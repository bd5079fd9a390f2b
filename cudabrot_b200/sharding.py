"""Host-side sharding of the sample stream over GPUs and the one-collective merge.

The reference is single-GPU (cudabrot.cu has no NCCL/MPI).  Samples are independent and the
histogram is a sum of integer increments, so GPU g of G renders a disjoint range of Philox sample
indices into a private histogram and one reduce(sum) to rank 0 merges them.
"""


def contiguous_split(first, count, rank, world):
    """Strong-scaling split of [first, first+count): rank r gets a contiguous block; the union over
    ranks is the whole range, so the merged histogram equals the 1-GPU histogram bit for bit.
    Returns (first_r, count_r)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per, extra = divmod(count, world)
    lo = per * rank + min(rank, extra)
    return first + lo, per + (1 if rank < extra else 0)


def step_range(step, rank, world, per_gpu, first=0):
    """Weak-scaling schedule used by bench.py: every (step, rank) pair owns a fresh block of
    `per_gpu` sample indices, laid out so that step s of all ranks is one contiguous range."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return first + (step * world + rank) * per_gpu, per_gpu


def merge_to_root(hist_tensor, root=0):
    """The single collective of the path: sum every rank's histogram into `root` in place
    (NCCL ncclReduce over NVLink for CUDA tensors, gloo for the CPU tests).  uint32 counters are
    carried as int32: two's-complement addition is the same bits, wrap-around included."""
    import torch
    import torch.distributed as dist
    if hist_tensor.dtype != torch.int32:
        raise TypeError("histogram must be viewed as int32 for the collective")
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(hist_tensor, dst=root, op=dist.ReduceOp.SUM)
    return hist_tensor
