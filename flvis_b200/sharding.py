"""Multi-GPU data plane for independent sequences.

Partition: stream i of a job with R ranks and S slots per rank lives on rank i // S, slot i % S.  Sequences never interact,
so the frame loop has NO collective; the only traffic between ranks is
  * scatter_inputs  -- optional, once before a run: rank 0 owns the input frames of every sequence and deals each rank its
                       block (grouped point-to-point sends; over NCCL these ride NVLink), for deployments where one process
                       ingests all cameras; the default is that every rank loads / renders its own sequences;
  * ResultGather    -- once per block of K frames: the fixed-size per-frame result records of every rank (pose 7 x f64 +
                       landmark count, optionally the landmark pixel positions and ids) are all-gathered on a side stream;
  * max_over_ranks  -- the timing reduction of the bench contract.
Everything goes through torch.distributed, so the same code runs on NCCL (GPU) and gloo (the CPU tests)."""
import torch

RESULT_WIDTH = 8          # per sequence and frame: T_c_w as [qx qy qz qw tx ty tz] + landmark count


def stream_ids(rank, world, streams_per_rank):
    """Global stream ids owned by `rank` (contiguous block, weak scaling: every rank owns streams_per_rank)."""
    return list(range(rank * streams_per_rank, (rank + 1) * streams_per_rank))


def stream_seed(global_stream_id, base=1000):
    """Seed of the synthetic sequence of a global stream id (independent of how streams are sharded)."""
    return base + global_stream_id


def owner_of(global_stream_id, streams_per_rank):
    """(rank, slot) of a global stream id."""
    return global_stream_id // streams_per_rank, global_stream_id % streams_per_rank


def max_over_ranks(value_ms, dist=None, device=None):
    """Job time = slowest rank (the bench contract); identity when not distributed."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value_ms)
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


def scatter_inputs(full, streams_per_rank, dist, device=None, stream_dim=1):
    """Rank 0 holds `full` = the inputs of ALL sequences (stream axis `stream_dim`, length world * streams_per_rank); every rank
    returns its own block (same dtype, stream axis of length streams_per_rank).  Other ranks pass full=None plus nothing else:
    shape and dtype travel first.  Grouped isend / irecv, one message per destination rank."""
    world, rank = dist.get_world_size(), dist.get_rank()
    meta = [None]
    if rank == 0:
        shape = list(full.shape); shape[stream_dim] = streams_per_rank
        meta = [(shape, str(full.dtype).replace("torch.", ""))]
    dist.broadcast_object_list(meta, src=0)
    shape, dtype = meta[0][0], getattr(torch, meta[0][1])
    if rank == 0:
        assert full.shape[stream_dim] == world * streams_per_rank
        reqs, keep = [], []
        for r in range(1, world):
            blk = full.narrow(stream_dim, r * streams_per_rank, streams_per_rank).contiguous()
            if device is not None:
                blk = blk.to(device, non_blocking=True)
            keep.append(blk)
            reqs.append(dist.isend(blk, dst=r))
        mine = full.narrow(stream_dim, 0, streams_per_rank).contiguous()
        if device is not None:
            mine = mine.to(device)
        for q in reqs:
            q.wait()
        return mine
    buf = torch.empty(shape, dtype=dtype, device=device)
    dist.irecv(buf, src=0).wait()
    return buf


class ResultGather:
    """Per-frame result records of K frames x S sequences per rank, all-gathered to every rank in ONE collective per block.

    record(frame_in_block, slot, pose7, n_landmarks) fills the rank's pinned staging block; flush() copies it to the device
    (side stream on CUDA) and starts all_gather_into_tensor; wait() makes the result visible and returns the
    [world][K][S][RESULT_WIDTH] tensor.  Nothing here is called inside a frame."""

    def __init__(self, dist, block_frames, streams_per_rank, device=None, compute_stream=None):
        self.dist, self.K, self.S, self.device = dist, block_frames, streams_per_rank, device
        self.world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
        cuda = device is not None and torch.device(device).type == "cuda"
        self.h = torch.zeros(block_frames, streams_per_rank, RESULT_WIDTH, dtype=torch.float64)
        if cuda:
            self.h = self.h.pin_memory()
        self.d = torch.zeros_like(self.h, device=device) if cuda else self.h
        self.all = torch.zeros(self.world, block_frames, streams_per_rank, RESULT_WIDTH, dtype=torch.float64, device=device if cuda else None)
        self.side = torch.cuda.Stream(device) if cuda else None
        self.compute_stream = compute_stream
        self.work = None
        self.hn = self.h.numpy()

    def record(self, frame_in_block, slot, pose7, n_landmarks):
        row = self.hn[frame_in_block, slot]
        row[:7] = pose7
        row[7] = n_landmarks

    def flush(self):
        if self.world == 1:
            self.all[0].copy_(self.h)
            return
        if self.side is not None:
            if self.compute_stream is not None:
                self.side.wait_stream(self.compute_stream)
            with torch.cuda.stream(self.side):
                self.d.copy_(self.h, non_blocking=True)
                self.work = self.dist.all_gather_into_tensor(self.all.view(-1), self.d.view(-1), async_op=True)
        else:
            self.work = self.dist.all_gather_into_tensor(self.all.view(-1), self.d.view(-1), async_op=True)

    def wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
        if self.side is not None and self.compute_stream is not None:
            self.compute_stream.wait_stream(self.side)
        return self.all

    def trajectory(self, global_stream_id):
        """[K][RESULT_WIDTH] records of one global sequence out of the gathered block."""
        r, s = owner_of(global_stream_id, self.S)
        return self.all[r, :, s]
