"""not gpu: GOP-sharded encoding (x264-vs2008_b200/gop_shard.py) — closed GOPs encoded by separate worker processes on two
gloo ranks and stitched by rank 0 must give, byte for byte, the stream ONE process of the unmodified reference writes.  Workers run the
performance-mode hooks against the CPU stand-in device (oracle/_ref/x264_b200_stub); tests/test_gpu_encode.py::test_gop_sharded_on_device
repeats it with the real library on the GPU."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_integration_host import _clip, _run, REF, STUB, _load_pkg  # noqa: E402

W, H, N, K = 96, 64, 23, 6   # 4 GOPs, the last one short
OPTS = "--qp 26 --me esa --merange 8 --subme 4 --bframes 2 --b-adapt 2 --ref 2"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_plan_and_fixups():
    _load_pkg()
    from x264_vs2008_b200 import gop_shard as G
    assert G.plan_gops(23, 6) == [(0, 0, 6), (1, 6, 6), (2, 12, 6), (3, 18, 5)]
    assert G.plan_gops(12, 6) == [(0, 0, 6), (1, 6, 6)]
    assert [G.gops_of_rank(7, 3, r) for r in range(3)] == [[0, 1, 2], [3, 4], [5, 6]]
    assert G.split_runs(G.plan_gops(23, 6), 2) == [(0, 0, 12), (2, 12, 11)]
    assert G.split_runs(G.plan_gops(23, 6), 8) == G.plan_gops(23, 6)
    with pytest.raises(ValueError):
        G.gops_of_rank(4, 2, 2)
    sei = b"\x00\x00\x00\x01\x06\x05\x10abc\x80"
    rest = b"\x00\x00\x00\x01\x67\x42\x00\x00\x00\x01\x68\xce"
    assert G.drop_leading_sei(sei + rest) == rest
    assert G.drop_leading_sei(rest) == rest
    assert G.stitch({1: sei + rest, 0: sei + rest}) == sei + rest + rest


def _worker(rank, world, port, src, tmp, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist
        _load_pkg()
        from x264_vs2008_b200 import gop_shard as G
        dist.init_process_group("gloo", rank=rank, world_size=world)
        gops = G.plan_gops(N, K)
        mine = [gops[k] for k in G.gops_of_rank(len(gops), world, rank)]
        parts, _ = G.encode_gops(STUB, src, W, H, OPTS.split(), K, mine if rank else G.split_runs(mine, 1), os.path.join(tmp, "r%d" % rank), workers=2)
        if not rank:   # rank 0 encoded its two GOPs as ONE run (one encoder invocation): hand it on under its first GOP's index
            parts = {0: parts[0], 1: b""}
        stream = G.gather_stream(dist, parts, len(gops))
        dist.barrier()
        dist.destroy_process_group()
        q.put((rank, sorted(parts), stream))
    except BaseException as e:
        q.put(("error", rank, repr(e)))
        raise


@pytest.mark.timeout(600)
def test_two_ranks_stitch_equals_single_process(tmp_path):
    if not (os.path.exists(REF) and os.path.exists(STUB)):
        pytest.skip("oracle/_ref builds not present (they are produced where the reference sources exist)")
    _load_pkg()
    from x264_vs2008_b200 import gop_shard as G
    src = str(tmp_path / "in.yuv")
    _clip(W, H, N, src)
    single = str(tmp_path / "single.264")
    r = _run(REF, OPTS + " " + " ".join(G.gop_options(K)), src, single, W, H)
    assert r.returncode == 0, r.stderr[-1500:]
    want = open(single, "rb").read()
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    port, world = _free_port(), 2
    procs = [mpc.Process(target=_worker, args=(rk, world, port, src, str(tmp_path), q)) for rk in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=500) for _ in procs]
    assert not any(g[0] == "error" for g in got), got
    got.sort(key=lambda g: g[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert got[0][1] == [0, 1] and got[1][1] == [2, 3]
    assert got[1][2] is None
    assert got[0][2] == want, (len(got[0][2]), len(want))
    # and in one process, four workers at once
    parts, _ = G.encode_gops(STUB, src, W, H, OPTS.split(), K, G.plan_gops(N, K), str(tmp_path / "solo"), workers=4)
    assert G.stitch(parts) == want
    # three workers, runs of 2 + 1 + 1 GOPs
    parts, _ = G.encode_gops(STUB, src, W, H, OPTS.split(), K, G.split_runs(G.plan_gops(N, K), 3), str(tmp_path / "runs"), workers=3)
    assert sorted(parts) == [0, 2, 3] and G.stitch(parts) == want


GOPS_STUB = os.path.join(ROOT, "oracle", "_ref", "x264_b200_gops_stub")


@pytest.mark.parametrize("workers", [1, 3, 4])
def test_in_process_front_end_equals_single_process(tmp_path, workers):
    """integration/x264_b200_gops.c — several encoder instances as threads of ONE process (per-thread hook state, cost tables built before
    the threads start, seeds handed to each instance's x264_encoder_open) — against one process of the unmodified reference"""
    import subprocess
    if not (os.path.exists(REF) and os.path.exists(GOPS_STUB)):
        pytest.skip("oracle/_ref builds not present (they are produced where the reference sources exist)")
    _load_pkg()
    from x264_vs2008_b200 import gop_shard as G
    src = str(tmp_path / "in.yuv")
    _clip(W, H, N, src)
    single, out = str(tmp_path / "single.264"), str(tmp_path / "gops.264")
    r = _run(REF, OPTS + " " + " ".join(G.gop_options(K)), src, single, W, H)
    assert r.returncode == 0, r.stderr[-1500:]
    r = subprocess.run([GOPS_STUB, "--no-asm"] + OPTS.split() + ["--keyint", str(K), "--workers", str(workers), "-o", out, src, "%dx%d" % (W, H)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    assert open(out, "rb").read() == open(single, "rb").read()
