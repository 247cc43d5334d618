"""world_size-2 gloo test (CPU) of the N>1 path: channel sharding, the equal-sized block gather and
the merge on rank 0 — the same `_shard` functions bench.py runs over NCCL on the GPU box."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_CH, N_FIELDS, CAP = 7, 39, 16


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _channel_block(c):
    """what the device would hold for channel c: [N_FIELDS][CAP] doubles"""
    return (c * 1000 + np.arange(N_FIELDS * CAP, dtype=np.float64)).reshape(N_FIELDS, CAP)


def _worker(rank, world, port, out_path):
    import sys
    sys.path.insert(0, ROOT)
    import bds3_b200  # noqa: F401
    from bds3_b200 import _shard
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _shard.shard_indices(N_CH, rank, world)
    block = torch.from_numpy(np.concatenate([_channel_block(c).reshape(-1) for c in mine]))
    pad = _shard.max_block_elems(block.numel(), dist)
    assert pad == 4 * N_FIELDS * CAP                      # rank 0 owns 4 of the 7 channels
    got = _shard.gather_blocks(block, pad, dist, dst=0)
    # acquisition: disjoint PRN shards, merged by summation
    lo, hi = _shard.prn_range(10, rank, world)
    part = np.zeros((3, 10))
    part[:, lo:hi] = rank + 1
    parts = [torch.zeros(30, dtype=torch.float64) for _ in range(world)] if rank == 0 else None
    dist.gather(torch.from_numpy(part.reshape(-1)), parts, dst=0)
    if rank == 0:
        merged = _shard.merge_channel_blocks(got, N_CH, N_FIELDS, CAP)
        acq = _shard.merge_acq_results(parts, 10)
        np.savez(out_path, merged=merged, acq=acq)
    else:
        assert got is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_gather_merge_world2(tmp_path):
    out = str(tmp_path / "merged.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    z = np.load(out)
    for c in range(N_CH):
        np.testing.assert_array_equal(z["merged"][c], _channel_block(c))
    np.testing.assert_array_equal(z["acq"][0], [1, 1, 1, 1, 1, 2, 2, 2, 2, 2])


def test_single_process_path_needs_no_process_group():
    import sys
    sys.path.insert(0, ROOT)
    import bds3_b200  # noqa: F401
    from bds3_b200 import _shard
    b = torch.arange(8, dtype=torch.float64)
    assert _shard.max_block_elems(8, None) == 8
    assert _shard.gather_blocks(b, 8, None)[0] is b
