"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: batch partitioning, the one
init-time table broadcast, and the max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phastft_b200.sharding import broadcast_blob, max_over_ranks, shard_range


def test_shard_ranges_partition_the_batch():
    for batch in (1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(batch, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == batch
            for (a0, a1), (b0, b1) in zip(ranges, ranges[1:]):
                assert a1 == b0 and a0 <= a1
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert [shard_range(4096, r, 8) for r in range(8)] == [(512 * r, 512 * (r + 1)) for r in range(8)]
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # rank 0 "builds" the planner tables; the others start from garbage
        blob = torch.arange(4096, dtype=torch.int64).to(torch.uint8) if rank == 0 else torch.full((4096,), 255, dtype=torch.uint8)
        broadcast_blob(blob, src=0)
        ok_blob = bool(torch.equal(blob, torch.arange(4096, dtype=torch.int64).to(torch.uint8)))
        # each rank transforms its shard of a seeded batch (numpy stand-in for the device kernels: this
        # test is about which transforms a rank owns, not about the kernels)
        batch, n = 10, 64
        data = np.random.default_rng(1234).standard_normal((batch, n)) + 0j
        lo, hi = shard_range(batch, rank, world)
        mine = np.fft.fft(data[lo:hi], axis=1)
        gathered = [None] * world
        dist.all_gather_object(gathered, (lo, hi, mine))
        full = np.concatenate([g[2] for g in sorted(gathered, key=lambda g: g[0])])
        ok_fft = bool(np.array_equal(full, np.fft.fft(data, axis=1)))
        t = max_over_ranks(1.0 + rank)
        out.put((rank, ok_blob, ok_fft, t))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_broadcast_and_shards():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_blob, ok_fft, t in res:
        assert ok_blob and ok_fft and t == 2.0
