"""GOP-sharded bit-identical encoding: closed GOPs are independent units of work.

With a fixed IDR cadence (--keyint K --min-keyint K --scenecut -1), constant QP, and the defaults --nr 0 / --direct spatial, nothing on
the data path of the reference crosses an IDR: reference lists, POC and frame_num restart (S/encoder/encoder.c:1094-1110, 1480-1484) and
the slice-type analysis never looks past the keyint limit (S/encoder/slicetype.c:494-495).  So GOP k = input frames [k*K, (k+1)*K) can be
encoded by any worker started like `x264 --seek k*K --frames K` (S/x264.c:519-524, 829), and the single-process stream is the
concatenation of the shards in GOP order after two fix-ups:
  * the IDR slices carry idr_pic_id = GOP index mod 65536 (encoder.c:1107-1110): the worker is told its first value through
    X264_B200_IDR_PIC_ID (integration/x264_b200_hooks.c seeds h->i_idr_pic_id after x264_encoder_open);
  * the CABAC flush of every slice embeds one bit of a signature selected by the count of frames coded so far (h->i_frame,
    S/common/cabac.c:917): the worker starts that count at its first frame number (X264_B200_CODED_FRAMES);
  * every worker starts its stream with the version SEI the reference writes for frame 0 only (encoder.c:1570-1578): dropped from
    every shard but the first.
A worker takes a RUN of consecutive GOPs in one encoder invocation (one CUDA context start-up per worker, not per GOP): ranks own
contiguous runs of the GOP list, and within a rank several worker processes may share the GPU (the per-frame device work is ~1 ms, the
sequential macroblock loop on the host ~60-100 ms), which is how one B200 feeds all host cores.
No collective on the data path; rank 0 gathers the NAL bytes.
"""
import os
import subprocess
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def plan_gops(n_frames, keyint):
    """[(gop index, first frame, frame count)]"""
    return [(k, s, min(keyint, n_frames - s)) for k, s in enumerate(range(0, n_frames, keyint))]


def split_runs(gops, n_parts):
    """consecutive GOPs -> at most n_parts runs of consecutive GOPs, sizes differing by at most one; a run is (first gop index, first
    frame, frame count) — the same shape as a GOP, so encode_gop() takes either"""
    n = len(gops)
    runs, q, r = [], n // max(1, n_parts), n % max(1, n_parts)
    lo = 0
    for p in range(min(n_parts, n)):
        hi = lo + q + (1 if p < r else 0)
        if hi > lo:
            runs.append((gops[lo][0], gops[lo][1], sum(g[2] for g in gops[lo:hi])))
        lo = hi
    return runs


def gops_of_rank(n_gops, world, rank):
    """contiguous [lo, hi) of the GOP list for `rank` (sizes differ by at most one)"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    q, r = divmod(n_gops, world)
    lo = rank * q + min(rank, r)
    return list(range(lo, lo + q + (1 if rank < r else 0)))


def gop_options(keyint):
    return ["--keyint", str(keyint), "--min-keyint", str(keyint), "--scenecut", "-1"]


def drop_leading_sei(data):
    """remove the first NAL if it is an SEI (type 6); x264 writes 4-byte start codes for every NAL (S/common/common.c:656-698)"""
    if data[:4] == b"\x00\x00\x00\x01" and (data[4] & 0x1f) == 6:
        nxt = data.find(b"\x00\x00\x00\x01", 4)
        if nxt > 0:
            return data[nxt:]
    return data


def encode_gop(exe, src, width, height, opts, keyint, gop, out_path, env=None, threads=1):
    """one worker: encode GOP (or run of GOPs) `gop` = (index of its first GOP, first frame, frame count) of `src` to out_path;
    returns (stderr, wall seconds)"""
    k, first, count = gop
    cmd = [exe, "--no-asm", "--threads", str(threads)] + list(opts) + gop_options(keyint) + ["--seek", str(first), "--frames", str(count), "-o", out_path, src,
                                                                                              "%dx%d" % (width, height)]
    e = dict(os.environ)
    e.update(env or {})
    e["X264_B200_IDR_PIC_ID"] = str(k % 65536)
    e["X264_B200_CODED_FRAMES"] = str(first)   # closed GOPs: every earlier frame has been coded when this IDR comes up
    t = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    if r.returncode != 0:
        raise RuntimeError("GOP %d: %s failed (%d): %s" % (k, exe, r.returncode, r.stderr[-800:]))
    return r.stderr, time.perf_counter() - t


def encode_gops(exe, src, width, height, opts, keyint, gops, tmp_dir, workers=1, env=None, tag="gop"):
    """encode the given GOPs with `workers` concurrent worker processes; -> {gop index: bytes}, wall seconds"""
    os.makedirs(tmp_dir, exist_ok=True)
    t = time.perf_counter()

    def one(g):
        path = os.path.join(tmp_dir, "%s_%05d.264" % (tag, g[0]))
        encode_gop(exe, src, width, height, opts, keyint, g, path, env)
        with open(path, "rb") as f:
            data = f.read()
        os.unlink(path)
        return g[0], data
    with ThreadPoolExecutor(max_workers=max(1, workers)) as pool:
        out = dict(pool.map(one, gops))
    return out, time.perf_counter() - t


def stitch(parts):
    """{gop index: bytes} of ALL GOPs -> the stream a single process would have written"""
    return b"".join(parts[k] if k == 0 else drop_leading_sei(parts[k]) for k in sorted(parts))


def gather_stream(dist, local_parts, n_gops):
    """every rank passes {gop: bytes}; rank 0 returns the stitched stream (others None).  Fixed-size uint8 tensors (lengths first), so it
    works on gloo and nccl alike."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return stitch(local_parts)
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    lens = torch.zeros(n_gops, dtype=torch.int64, device=dev)
    for k, b in local_parts.items():
        lens[k] = len(b)
    dist.all_reduce(lens, op=dist.ReduceOp.SUM)          # each GOP is owned by exactly one rank
    per_rank = [int(sum(int(lens[k]) for k in gops_of_rank(n_gops, world, r))) for r in range(world)]
    cap = max(per_rank + [1])
    mine = np.zeros(cap, np.uint8)
    blob = b"".join(local_parts[k] for k in gops_of_rank(n_gops, world, rank))
    mine[:len(blob)] = np.frombuffer(blob, np.uint8)
    outs = [torch.empty(cap, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(outs, torch.from_numpy(mine).to(dev))
    if rank != 0:
        return None
    parts = {}
    for r in range(world):
        raw, off = outs[r].cpu().numpy().tobytes(), 0
        for k in gops_of_rank(n_gops, world, r):
            parts[k] = raw[off:off + int(lens[k])]
            off += int(lens[k])
    return stitch(parts)
