"""Multi-GPU host logic: frames (or frame pairs) shard across ranks, one process per GPU, no data-path collective.

The reference's own parallelism is frame-level too (sliceless threads, S/encoder/encoder.c:1569-1608: one x264_t per
frame in flight); here a rank owns a contiguous run of frame pairs — contiguous so that a reference frame is uploaded to
the GPU that also searches against it.  The only cross-rank traffic is the end-of-job reduction of timings and counters
(torch.distributed, gloo on CPU in the tests, nccl on the GPUs), and optionally a gather of per-frame results to rank 0.
"""
import numpy as np


def frame_shard(n_units, world, rank):
    """contiguous [lo, hi) of `n_units` for `rank`; sizes differ by at most one, earlier ranks take the remainder"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    q, r = divmod(n_units, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def reduce_job(dist, device, elapsed_ms, counters):
    """(max over ranks of each elapsed_ms entry, sum over ranks of each counter) — the bench contract's timing rule.
    `dist` is torch.distributed or None (single process)."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(x) for x in elapsed_ms], [float(x) for x in counters]
    t = torch.tensor(list(elapsed_ms), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    c = torch.tensor(list(counters), dtype=torch.float64, device=device)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t], [float(x) for x in c]


def gather_results(dist, local, n_units, itemshape, dtype):
    """every rank passes the results of its shard (array [hi-lo, *itemshape]); rank 0 gets the whole job's array in unit order
    (others get None).  Uses all_gather_object-free fixed-size tensors so it works the same on gloo and nccl."""
    import torch
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(local, dtype)
    world, rank = dist.get_world_size(), dist.get_rank()
    per = -(-n_units // world)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = np.zeros((per,) + tuple(itemshape), dtype)
    buf[:len(local)] = local
    mine = torch.from_numpy(buf.view(np.uint8).reshape(-1)).to(dev)
    outs = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(outs, mine)
    if rank != 0:
        return None
    parts = []
    for r in range(world):
        lo, hi = frame_shard(n_units, world, r)
        a = outs[r].cpu().numpy().view(dtype).reshape((per,) + tuple(itemshape))
        parts.append(a[:hi - lo])
    return np.concatenate(parts)
