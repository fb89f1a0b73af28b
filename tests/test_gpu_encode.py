"""gpu: the performance-mode integration (integration/x264_b200_hooks.c linked with the unmodified reference and the real libx264_cuda.so:
integration/_build/x264_b200) must write byte-identical bitstreams to the plain reference CLI while the exhaustive search reads device
grids and deblocking / half-pel planes come from the device (BASELINE.json metric "bit-exact encode")."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_integration_host import _clip, _run, REF  # noqa: E402

B200 = os.path.join(ROOT, "integration", "_build", "x264_b200")

CONFIGS = [
    ("esa_subme2_cif", 352, 288, 5, "--me esa --merange 16 --subme 2", {}),
    ("esa_subme1_slack2", 96, 64, 3, "--me esa --merange 8 --subme 1", {"X264_B200_GRID_SLACK": "2"}),
    ("esa_subme5_chroma_me", 176, 144, 4, "--me esa --merange 12 --subme 5 --8x8dct", {}),
    ("esa_subme7_rd_refs", 176, 144, 5, "--me esa --merange 8 --subme 7 --8x8dct --ref 3 --mixed-refs", {}),
    ("esa_b_frames", 176, 144, 7, "--me esa --merange 8 --subme 4 --bframes 2 --b-adapt 2 --weightb --ref 2", {}),
    ("esa_p4x4", 64, 48, 3, "--me esa --merange 8 --subme 2 --partitions all", {}),
    ("esa_nodeblock", 96, 64, 3, "--me esa --merange 8 --subme 2 --no-deblock", {}),
    ("esa_cavlc_deblock_offsets", 96, 64, 3, "--me esa --merange 8 --subme 3 --no-cabac --8x8dct --deblock 2:-1", {}),
    ("esa_crf_aq", 176, 144, 5, "--crf 24 --me esa --merange 8 --subme 6 --bframes 1", {}),
    ("esa_odd_size", 100, 60, 3, "--me esa --merange 16 --subme 2", {}),
    ("esa_merange32", 176, 144, 3, "--me esa --merange 32 --subme 2", {}),
    ("hex_frame_end_only", 176, 144, 4, "--me hex --subme 5 --bframes 1", {}),
    ("tesa_left_to_reference", 64, 48, 3, "--me tesa --merange 8 --subme 4", {}),
    ("esa_1080p", 1920, 1080, 3, "--me esa --merange 16 --subme 2", {}),                                   # BASELINE config 2, end to end
    ("esa_1080p_subme7_8x8dct", 1920, 1080, 2, "--me esa --merange 16 --subme 7 --8x8dct", {}),            # BASELINE config 3
]


@pytest.mark.parametrize("tag,w,h,n,opts,env", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_encode_bitstream_identical(tmp_path, tag, w, h, n, opts, env):
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("oracle/_ref/x264 or integration/_build/x264_b200 not present (they are built where the reference sources exist)")
    src = str(tmp_path / "in.yuv")
    _clip(w, h, n, src)
    a, b = str(tmp_path / "ref.264"), str(tmp_path / "b200.264")
    r0 = _run(REF, opts, src, a, w, h)
    assert r0.returncode == 0, r0.stderr[-2000:]
    e = {"X264_B200_VERBOSE": "1"}
    e.update(env)
    r1 = _run(B200, opts, src, b, w, h, e)
    assert r1.returncode == 0, r1.stderr[-2000:]
    assert open(a, "rb").read() == open(b, "rb").read(), "bitstreams differ\n" + r1.stderr[-1500:]
    stat = lambda s: [l for l in s.splitlines() if re.search(r"PSNR Mean|SSIM Mean|x264 \[info\]: slice", l)]
    assert stat(r0.stderr) == stat(r1.stderr)
    m = re.search(r"x264_b200: (\d+) ESA searches read device grids", r1.stderr)
    assert m, r1.stderr[-1500:]
    if "--me esa" in opts:
        assert int(m.group(1)) > 0
        assert "0 searches left to the reference" in r1.stderr
    k = re.search(r"(\d+) end-of-frame device passes; (\d+) kernel launches", r1.stderr)
    assert k and int(k.group(1)) > 0 and int(k.group(2)) > 0
    print(tag, r1.stderr.strip().splitlines()[-2:])


def test_gop_sharded_on_device(tmp_path):
    """GOP-sharded encoding with the real library: four worker processes share cuda:0, the stitched stream equals ONE process of the
    unmodified reference (x264-vs2008_b200/gop_shard.py; the two-rank gloo variant runs on the CPU in tests/test_gop_shard.py)"""
    if not (os.path.exists(REF) and os.path.exists(B200)):
        pytest.skip("builds not present")
    from test_integration_host import _load_pkg
    _load_pkg()
    from x264_vs2008_b200 import gop_shard as G
    w, h, n, k = 352, 288, 22, 6
    opts = "--qp 26 --me esa --merange 16 --subme 5 --bframes 2 --b-adapt 2 --ref 2"
    src = str(tmp_path / "in.yuv")
    _clip(w, h, n, src)
    single = str(tmp_path / "single.264")
    r = _run(REF, opts + " " + " ".join(G.gop_options(k)), src, single, w, h)
    assert r.returncode == 0, r.stderr[-1500:]
    parts, wall = G.encode_gops(B200, src, w, h, opts.split(), k, G.plan_gops(n, k), str(tmp_path / "shards"), workers=4)
    assert G.stitch(parts) == open(single, "rb").read()


@pytest.mark.parametrize("size,n,k,workers,opts", [((352, 288), 30, 6, 5, "--qp 26 --me esa --merange 16 --subme 5 --bframes 2 --b-adapt 2 --ref 2"),
                                                   ((1920, 1080), 12, 4, 3, "--qp 26 --me esa --merange 16 --subme 2 --no-psnr --no-ssim")])
def test_gop_parallel_front_end_on_device(tmp_path, size, n, k, workers, opts):
    """integration/x264_b200_gops: encoder threads of one process sharing cuda:0 (one device context and stream per thread) write, stitched,
    the stream of one reference process"""
    gops = os.path.join(ROOT, "integration", "_build", "x264_b200_gops")
    if not (os.path.exists(REF) and os.path.exists(gops)):
        pytest.skip("builds not present")
    from test_integration_host import _load_pkg
    _load_pkg()
    from x264_vs2008_b200 import gop_shard as G
    w, h = size
    src = str(tmp_path / "in.yuv")
    _clip(w, h, n, src)
    single, out = str(tmp_path / "single.264"), str(tmp_path / "gops.264")
    r = _run(REF, opts + " " + " ".join(G.gop_options(k)), src, single, w, h)
    assert r.returncode == 0, r.stderr[-1500:]
    e = dict(os.environ)
    e["X264_B200_VERBOSE"] = "1"
    r = subprocess.run([gops, "--no-asm"] + opts.split() + ["--keyint", str(k), "--workers", str(workers), "-o", out, src, "%dx%d" % (w, h)],
                       capture_output=True, text=True, timeout=900, env=e)
    assert r.returncode == 0, r.stderr[-1500:]
    assert open(out, "rb").read() == open(single, "rb").read()
    assert len(re.findall(r"ESA searches read device grids", r.stderr)) == workers
