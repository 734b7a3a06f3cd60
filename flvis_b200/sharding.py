"""Multi-GPU sharding of independent sequences: stream i of a job with R ranks and S slots per rank
lives on rank i // S, slot i % S; there is no data-path collective -- the only cross-rank traffic is the barrier and
the max-over-ranks of the device time, plus an optional gather of per-rank result digests."""


def stream_ids(rank, world, streams_per_rank):
    """Global stream ids owned by `rank` (contiguous block, weak scaling: every rank owns streams_per_rank)."""
    return list(range(rank * streams_per_rank, (rank + 1) * streams_per_rank))


def stream_seed(global_stream_id, base=1000):
    """Seed of the synthetic sequence of a global stream id (independent of how streams are sharded)."""
    return base + global_stream_id


def max_over_ranks(value_ms, dist=None, device=None):
    """Job time = slowest rank (the bench contract); identity when not distributed."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value_ms)
    import torch
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_digests(digest, dist=None):
    """All ranks' per-rank result digests (python ints) on every rank, rank order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [digest]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, digest)
    return out
