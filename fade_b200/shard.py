"""Multi-GPU plumbing of the path: reads are independent units, so ranks shard the
read range, every GPU holds its own reference copy, and there is NO data-path collective -- only
the timing / counter reductions below (NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import os


def world():
    """(rank, local_rank, world_size) from the torchrun environment (1 process when absent)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_range(total: int, rank: int, world_size: int) -> tuple[int, int]:
    """Contiguous, balanced [first, last) slice of `total` reads for `rank` (strong scaling)."""
    base, rem = divmod(total, world_size)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def weak_range(per_rank: int, rank: int) -> tuple[int, int]:
    """Weak scaling: every rank simulates its own `per_rank` reads of the global read stream."""
    return rank * per_rank, (rank + 1) * per_rank


class Reducer:
    """max / sum over ranks of host scalars; a no-op in a single process."""

    def __init__(self, world_size: int, device=None):
        self.world_size = world_size
        self.device = device

    def _reduce(self, x: float, op: str) -> float:
        if self.world_size == 1:
            return float(x)
        import torch
        import torch.distributed as dist
        t = torch.tensor([float(x)], dtype=torch.float64, device=self.device or "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
        return float(t.item())

    def max(self, x: float) -> float:
        return self._reduce(x, "max")

    def sum(self, x: float) -> float:
        return self._reduce(x, "sum")
