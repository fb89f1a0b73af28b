"""not-gpu: the N>1 host path (frame sharding, max/sum reduction, result gather) with 2 gloo ranks on CPU.
Each rank runs the ORACLE's full-pel search on its shard of frame pairs (standing in for the device call, which has no CPU
path); rank 0 checks that the gathered job equals the single-process job."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pair_result(pair):
    """oracle ESA over a handful of blocks of frame pair `pair` -> int32[n,3]"""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import __graft_entry__ as ge
    import xo_api as X
    from helpers import make_me_jobs, oracle_me
    pkg = ge.load_pkg()
    from x264_vs2008_b200 import synth
    o = X.port()
    w, h = 96, 64
    g = o.geometry(w, h)
    clip = synth.Clip(w, h, seed=3)
    pe, pr = o.plane_from_picture(g, clip.luma(pair + 1)), o.plane_from_picture(g, clip.luma(pair))
    _, mis = make_me_jobs(pkg, g, seed=pair, n=6, me_range=8, qp=26)
    outs = oracle_me(o, g, pe, pr, None, mis)
    return np.array(outs, np.int32)


def _worker(rank, world, port, n_pairs, q):
    try:
        _worker_body(rank, world, port, n_pairs, q)
    except BaseException as e:  # report instead of letting the parent wait for its queue timeout
        q.put(("error", rank, repr(e)))
        raise


def _worker_body(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_pkg()
    from x264_vs2008_b200 import shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = shard.frame_shard(n_pairs, world, rank)
    local = np.stack([_pair_result(p) for p in range(lo, hi)]) if hi > lo else np.zeros((0, 6, 3), np.int32)
    ms, cnt = shard.reduce_job(dist, "cpu", [10.0 + rank, 5.0 - rank], [hi - lo, 7])
    allr = shard.gather_results(dist, local, n_pairs, (6, 3), np.int32)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, lo, hi, ms, cnt, None if allr is None else allr.tolist()))


def test_frame_shard_partition():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    ge.load_pkg()
    from x264_vs2008_b200 import shard
    for n in (0, 1, 5, 8, 33):
        for world in (1, 2, 3, 8):
            spans = [shard.frame_shard(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.frame_shard(4, 2, 2)


@pytest.mark.timeout(300)
def test_two_ranks_gloo(pkg):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port, n_pairs, world = _free_port(), 5, 2
    procs = [mpc.Process(target=_worker, args=(r, world, port, n_pairs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=240) for _ in procs]
    assert not any(g[0] == "error" for g in got), got
    got.sort()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert [(g[1], g[2]) for g in got] == [(0, 3), (3, 5)]
    for g in got:  # max of times, sum of counters, identical on both ranks
        assert g[3] == [11.0, 5.0] and g[4] == [5.0, 14.0]
    assert got[1][5] is None
    want = np.stack([_pair_result(p) for p in range(n_pairs)])
    assert np.array_equal(np.array(got[0][5], np.int32), want)
