"""N > 1 host logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ml_accelerated_simulation_b200.ensemble import aggregate_throughput, max_over_ranks, shard_batch


def test_shard_batch_partitions_exactly():
    for total in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_batch(total, r, world) for r in range(world)]
            assert blocks[0][0] == 0
            assert sum(c for _, c in blocks) == total
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1                                   # contiguous, disjoint
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1                      # balanced
    assert shard_batch(4096, 3, 8) == (1536, 512)                      # BASELINE config 4 at 8 GPUs
    with pytest.raises(ValueError):
        shard_batch(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    start, count = shard_batch(65, rank, world)
    slow = max_over_ranks(1.0 + rank)                                   # rank 1 is the slow one
    rate = aggregate_throughput(units_this_rank=count * 100.0, seconds_this_rank=1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, (start, count))
    if rank == 0:
        torch.save({"slow": slow, "rate": rate, "blocks": gathered}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_timing_and_throughput(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["slow"] == 2.0                                           # max over ranks, never rank 0's own time
    assert res["blocks"] == [(0, 33), (33, 32)]
    assert res["rate"] == pytest.approx(65 * 100.0 / 2.0)               # all units / slowest rank
