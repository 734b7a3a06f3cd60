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


def _plane_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S, K = 3, 4
    # scatter: rank 0 owns the frames of all world*S sequences ([frames][streams][h][w]); every rank gets its block
    full = None
    if rank == 0:
        full = torch.arange(5 * world * S * 2 * 2, dtype=torch.uint8).reshape(5, world * S, 2, 2)
    mine = sharding.scatter_inputs(full, S, dist)
    # gather: every rank records K frames x S sequences; all ranks see everything
    g = sharding.ResultGather(dist, K, S)
    for k in range(K):
        for s in range(S):
            gid = rank * S + s
            g.record(k, s, [0, 0, 0, 1, gid, k, 0.5], 100 + gid)
    g.flush()
    allr = g.wait().clone()
    one = g.trajectory(world * S - 1).clone()
    dist.barrier()
    q.put((rank, mine.numpy().tolist(), allr.numpy().tolist(), one.numpy().tolist()))
    dist.destroy_process_group()


def test_two_rank_scatter_and_result_gather():
    import numpy as np
    world, port, S, K = 2, 29517, 3, 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_plane_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps: p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps: p.join(timeout=60)
    full = np.arange(5 * world * S * 2 * 2, dtype=np.uint8).reshape(5, world * S, 2, 2)
    for r in range(world):
        assert np.array_equal(np.array(res[r][1], np.uint8), full[:, r * S:(r + 1) * S])      # each rank got exactly its block
        a = np.array(res[r][2])
        assert a.shape == (world, K, S, sharding.RESULT_WIDTH)
        for rr in range(world):
            for k in range(K):
                for s in range(S):
                    gid = rr * S + s
                    assert a[rr, k, s, 4] == gid and a[rr, k, s, 5] == k and a[rr, k, s, 7] == 100 + gid
        assert np.array(res[r][3])[:, 4].tolist() == [world * S - 1] * K                      # trajectory() addresses by global id
    assert sharding.owner_of(4, 3) == (1, 1)
