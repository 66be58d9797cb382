"""Host-side logic for the batched multi-GPU path.

One process per GPU (torchrun).  Batched transforms are independent units: rank g owns the
contiguous range [g*B/G, (g+1)*B/G) of the batch-major planar arrays.  The only communication
is one broadcast of the planner-table blob at init (so every rank holds bit-identical tables);
there is no collective on the data path.
"""
from __future__ import annotations


def shard_range(batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range of transforms owned by `rank` -- the same partition
    phastft_fft_dit_*_batch_sharded_host uses inside one process."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return batch * rank // world, batch * (rank + 1) // world


def broadcast_blob(blob, src: int = 0, group=None):
    """Broadcast a uint8 tensor holding the planner tables from `src` to every rank, in place.
    Works with any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
    import torch.distributed as dist
    dist.broadcast(blob, src=src, group=group)
    return blob


def max_over_ranks(value: float, device=None, group=None) -> float:
    """Timing rule: a multi-GPU step takes as long as its slowest rank."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
