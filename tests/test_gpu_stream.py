"""gpu: stream-level parity.  The UNMODIFIED reference encoder, linked so that x264_{pixel,mc,dct,quant}_init put the CUDA
back-end's entries on top of the C tables (oracle/ref_cuda_shim.c, the hook INTEGRATION.md describes), must write the same bytes as the
plain C build — every table entry is then exercised by the reference's own callers, with the reference's own data."""
import os
import re
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "x264")
CUD = os.path.join(ROOT, "oracle", "_ref", "x264_cuda")

CONFIGS = [
    ("config1_dia", 80, 64, 3, "--me dia --subme 1"),
    ("hex_8x8dct_b", 80, 64, 4, "--me hex --subme 5 --8x8dct --bframes 1"),
    ("esa", 80, 64, 2, "--me esa --merange 8 --subme 2"),                                                 # config 2's search
    ("umh_rd_weightb", 64, 48, 5, "--me umh --subme 7 --8x8dct --bframes 2 --b-adapt 2 --weightb --mixed-refs --ref 2"),  # config 3
    ("tesa_b3", 64, 48, 5, "--me tesa --merange 8 --subme 6 --bframes 3 --b-adapt 2"),                    # config 4's options
    ("cqm_jvt_qcif", 176, 144, 2, "--me hex --subme 4 --cqm jvt"),
]


def _clip(pkg, w, h, n, path):
    from x264_vs2008_b200 import synth
    clip = synth.Clip(w, h, seed=3)
    with open(path, "wb") as f:
        for i in range(n):
            for p in clip.yuv420(i):
                f.write(np.ascontiguousarray(p).tobytes())


@pytest.mark.parametrize("tag,w,h,n,opts", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_bitstream_identical(pkg, ctx, tmp_path, tag, w, h, n, opts):
    if not (os.path.exists(REF) and os.path.exists(CUD)):
        pytest.skip("oracle/_ref CLI builds not present (they are produced where the reference sources exist)")
    src = str(tmp_path / "in.yuv")
    _clip(pkg, w, h, n, src)
    outs, launches = [], 0
    for exe in (REF, CUD):
        out = str(tmp_path / (os.path.basename(exe) + ".264"))
        r = subprocess.run([exe, "--qp", "26", "--no-asm", "--threads", "1"] + opts.split() + ["-o", out, src, "%dx%d" % (w, h)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "encoded %d frames" % n in r.stderr + r.stdout
        outs.append(open(out, "rb").read())
        m = re.search(r"ref_cuda_shim: (\d+) device launches", r.stderr)
        if m:
            launches = int(m.group(1))
    assert launches > 1000 * n, launches   # the table entries really ran on the device
    assert len(outs[0]) > 500 and outs[0] == outs[1], (tag, len(outs[0]), len(outs[1]))
