"""not gpu: the HOST logic of the performance-mode integration (integration/x264_b200_hooks.c — grid bookkeeping, predictor stage, raster
argmin over quadrant grids, sub-pel stage, deferred end-of-frame pass, deferred PSNR/SSIM) checked for byte-identical bitstreams against
the unmodified reference CLI.  The device entry points are served by a CPU stand-in built from the oracle (oracle/cuda_stub.c ->
oracle/_ref/x264_b200_stub); the same hooks linked against the real libx264_cuda.so are tested on the GPU by tests/test_gpu_encode.py."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "x264")
STUB = os.path.join(ROOT, "oracle", "_ref", "x264_b200_stub")

CONFIGS = [
    ("esa_subme2", 96, 64, 4, "--me esa --merange 16 --subme 2", {}),
    ("esa_subme1_slack2", 96, 64, 3, "--me esa --merange 8 --subme 1", {"X264_B200_GRID_SLACK": "2"}),   # tight grids: many recomputed
    ("esa_subme5_chroma_me", 96, 64, 4, "--me esa --merange 12 --subme 5 --8x8dct", {}),
    ("esa_subme7_rd_refs", 96, 64, 5, "--me esa --merange 8 --subme 7 --8x8dct --ref 3 --mixed-refs", {}),
    ("esa_b_frames", 96, 64, 7, "--me esa --merange 8 --subme 4 --bframes 2 --b-adapt 2 --weightb --ref 2", {}),
    ("esa_p4x4", 64, 48, 3, "--me esa --merange 8 --subme 2 --partitions all", {}),                      # sub-8x8: one-job device searches
    ("esa_psnr_ssim_nodeblock", 96, 64, 3, "--me esa --merange 8 --subme 2 --no-deblock", {}),
    ("esa_cavlc_deblock_offsets", 96, 64, 3, "--me esa --merange 8 --subme 3 --no-cabac --8x8dct --deblock 2:-1", {}),
    ("esa_crf_aq", 96, 64, 5, "--crf 24 --me esa --merange 8 --subme 6 --bframes 1", {}),
    ("esa_odd_size", 100, 60, 3, "--me esa --merange 16 --subme 2", {}),                                  # not a multiple of 16
    ("hex_frame_end_only", 96, 64, 4, "--me hex --subme 5 --bframes 1", {}),                             # only the end-of-frame pass is hooked
    ("tesa_left_to_reference", 64, 48, 3, "--me tesa --merange 8 --subme 4", {}),                        # host TESA on the device's integral image
]


def _load_pkg():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    return g.load_pkg()


def _clip(w, h, n, path):
    _load_pkg()
    from x264_vs2008_b200 import synth
    clip = synth.Clip(w, h, seed=5)
    with open(path, "wb") as f:
        for i in range(n):
            for p in clip.yuv420(i):
                f.write(np.ascontiguousarray(p).tobytes())


def _run(exe, opts, src, out, w, h, env=None):
    cmd = [exe, "--no-asm", "--threads", "1"] + ([] if "--crf" in opts else ["--qp", "26"]) + opts.split() + ["-o", out, src, "%dx%d" % (w, h)]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e)


@pytest.mark.parametrize("tag,w,h,n,opts,env", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_host_logic_bitstream_identical(tmp_path, tag, w, h, n, opts, env):
    if not (os.path.exists(REF) and os.path.exists(STUB)):
        pytest.skip("oracle/_ref builds not present (they are produced where the reference sources exist)")
    src = str(tmp_path / "in.yuv")
    _clip(w, h, n, src)
    a, b = str(tmp_path / "ref.264"), str(tmp_path / "b200.264")
    r0 = _run(REF, opts, src, a, w, h)
    assert r0.returncode == 0, r0.stderr[-2000:]
    e = {"X264_B200_VERBOSE": "1"}
    e.update(env)
    r1 = _run(STUB, opts, src, b, w, h, e)
    assert r1.returncode == 0, r1.stderr[-2000:]
    assert open(a, "rb").read() == open(b, "rb").read(), "bitstreams differ\n" + r1.stderr[-1500:]
    # the statistics lines (PSNR / SSIM per frame type) must agree as well: the deferred slabs add up to the same numbers
    stat = lambda s: [l for l in s.splitlines() if re.search(r"PSNR Mean|SSIM Mean|x264 \[info\]: slice", l)]
    assert stat(r0.stderr) == stat(r1.stderr)
    m = re.search(r"x264_b200: (\d+) ESA searches read device grids \((\d+) macroblock grids recomputed", r1.stderr)
    assert m, r1.stderr[-1500:]
    if "--me esa" in opts:
        assert int(m.group(1)) > 0
    assert re.search(r"(\d+) end-of-frame device passes", r1.stderr) and int(re.search(r"(\d+) end-of-frame device passes", r1.stderr).group(1)) > 0
