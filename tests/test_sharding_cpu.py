"""world_size-2 gloo test of the multi-GPU plumbing (no GPU): stream partition, max-over-ranks timing, digest gather."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from flvis_b200 import sharding


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = sharding.stream_ids(rank, world, 4)
    t = sharding.max_over_ranks(10.0 + 5.0 * rank, dist)
    seeds = [sharding.stream_seed(i) for i in ids]
    digests = sharding.gather_digests(sum(seeds), dist)
    dist.barrier()
    q.put((rank, ids, t, digests))
    dist.destroy_process_group()


def test_two_rank_partition_and_timing():
    world, port = 2, 29511
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps: p.join(timeout=60)
    all_ids = res[0][1] + res[1][1]
    assert all_ids == list(range(8))                       # disjoint, complete, rank-major
    assert res[0][2] == res[1][2] == 15.0                  # job time = slowest rank
    assert res[0][3] == res[1][3] == [sum(1000 + i for i in range(4)), sum(1000 + i for i in range(4, 8))]


def test_single_process_is_identity():
    assert sharding.stream_ids(0, 1, 3) == [0, 1, 2]
    assert sharding.max_over_ranks(3.5) == 3.5
    assert sharding.gather_digests(7) == [7]
