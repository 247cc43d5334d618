"""Multi-GPU plumbing: one process per GPU, channels (tracking) or PRNs (acquisition) sharded
across the ranks, one gather of the packed result blocks to rank 0.

Channels never interact (BDS-3_B1C/WB_tracking.m:162 is a plain serial ``for``) and PRNs never
interact (BDS-3_B1C/acquisition.m:169), so there is no data-path collective: the IF record is
replicated in every GPU's HBM and the only exchange is the result gather.  The functions take a
``torch.distributed`` module (or None for a single process) so the same code runs over NCCL on
the GPU box and over gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_indices(n_units: int, rank: int, world: int) -> list[int]:
    """Round-robin ownership: unit i belongs to rank i % world (channel c -> GPU c mod G)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_units, world))


def shard_list(units, rank: int, world: int):
    return [units[i] for i in shard_indices(len(units), rank, world)]


def prn_range(n_prn: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous PRN shard [lo, hi) of the acquisition list (63 -> 8,8,8,8,8,8,8,7): every rank owns
    whole PRN rows, which the B2a second-peak metric needs (BDS-3_B2a/acquisition.m:249)."""
    base, extra = divmod(n_prn, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def max_block_elems(local_elems: int, dist=None) -> int:
    """Largest per-rank block (ranks differ by at most one channel): the gather uses equal-sized blocks."""
    if dist is None or dist.get_world_size() == 1:
        return int(local_elems)
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(local_elems)], device=dev, dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return int(t.item())


def gather_blocks(block, pad_to: int, dist=None, dst: int = 0):
    """One gather of the per-rank packed result block (1-D float64 tensor) to ``dst``.
    Returns the list of per-rank blocks on ``dst`` (None elsewhere)."""
    import torch
    if dist is None or dist.get_world_size() == 1:
        return [block]
    n = block.numel()
    if n < pad_to:
        block = torch.cat([block, torch.zeros(pad_to - n, dtype=block.dtype, device=block.device)])
    rank, world = dist.get_rank(), dist.get_world_size()
    out = [torch.empty_like(block) for _ in range(world)] if rank == dst else None
    dist.gather(block, out, dst=dst)
    return out


def merge_channel_blocks(blocks, n_channels: int, n_fields: int, capacity: int) -> np.ndarray:
    """Undo the round-robin shard on rank 0: blocks[r] holds channels r, r+G, ... as
    [n_local][n_fields][capacity] doubles -> [n_channels][n_fields][capacity]."""
    world = len(blocks)
    out = np.empty((n_channels, n_fields, capacity))
    per = n_fields * capacity
    for r, b in enumerate(blocks):
        idx = shard_indices(n_channels, r, world)
        a = b.detach().cpu().numpy() if hasattr(b, "detach") else np.asarray(b)
        out[idx] = a[: len(idx) * per].reshape(len(idx), n_fields, capacity)
    return out


def merge_acq_results(parts, max_prn: int) -> np.ndarray:
    """parts[r] = [3][max_prn] (carrFreq, codePhase, peakMetric) with zeros outside rank r's PRN
    shard -> element-wise sum (shards are disjoint, zero = not found, acquisition.m:161-165)."""
    out = np.zeros((3, max_prn))
    for p in parts:
        a = p.detach().cpu().numpy() if hasattr(p, "detach") else np.asarray(p)
        out += a.reshape(3, max_prn)
    return out
