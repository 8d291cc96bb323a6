"""Index-range sharding across ranks (one process per GPU under torchrun).

The path has no exchange step: every point is independent, so rank k processes
the contiguous range [floor(k n / G), floor((k+1) n / G)) of each section and
writes its disjoint slice of the output.  The only cross-rank
values are scalars reduced on the host side of torch.distributed: the lowest
failing index (MIN), the step time (MAX) and counters (SUM).  With the nccl
backend the scalars travel as device tensors; with gloo (CPU tests) as host
tensors.  libptau_b200.so applies the same range formula inside one process when
a context owns several GPUs (capi.cu: ptau_convert).
"""
from __future__ import annotations

from typing import Tuple

STATUS_NONE = 0xFFFFFFFFFFFFFFFF
_I64_MAX = (1 << 63) - 1


def shard_range(n_points: int, rank: int, world: int) -> Tuple[int, int]:
    return (n_points * rank) // world, (n_points * (rank + 1)) // world


def _dist():
    import torch
    import torch.distributed as dist

    dev = "cpu"
    if dist.is_initialized() and dist.get_backend() == "nccl":
        dev = "cuda:%d" % torch.cuda.current_device()
    return torch, dist, dev


def reduce_status(local_status: int) -> int:
    """MIN over ranks of (index << 8 | kind); STATUS_NONE when every rank is clean."""
    torch, dist, dev = _dist()
    v = _I64_MAX if local_status == STATUS_NONE else int(local_status)
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([v], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        v = int(t.item())
    return STATUS_NONE if v == _I64_MAX else v


def reduce_max_ms(ms: float) -> float:
    torch, dist, dev = _dist()
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(ms)


def reduce_sum(v: float) -> float:
    torch, dist, dev = _dist()
    if dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        r = float(t.item())
        return int(r) if float(r).is_integer() else r
    return v
