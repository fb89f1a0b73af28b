"""not gpu: the renamed-symbol build of the reference encoder (oracle/_ref/x264_cuda) with every hook switched off must behave exactly
like the plain C build — the wrappers themselves change nothing.  (With the hooks on it needs a GPU: tests/test_gpu_stream.py.)"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "x264")
CUD = os.path.join(ROOT, "oracle", "_ref", "x264_cuda")


def test_shim_is_transparent_without_hooks(pkg, tmp_path):
    if not (os.path.exists(REF) and os.path.exists(CUD)):
        pytest.skip("oracle/_ref CLI builds not present")
    from x264_vs2008_b200 import synth
    w, h, n = 96, 64, 4
    clip = synth.Clip(w, h, seed=3)
    src = str(tmp_path / "in.yuv")
    with open(src, "wb") as f:
        for i in range(n):
            for p in clip.yuv420(i):
                f.write(np.ascontiguousarray(p).tobytes())
    outs = []
    for exe, env in ((REF, {}), (CUD, {"X264_CUDA_TABLES": "none", "X264_CUDA_FRAME_HOOKS": "0"})):
        out = str(tmp_path / (os.path.basename(exe) + ".264"))
        r = subprocess.run([exe, "--qp", "26", "--no-asm", "--threads", "1", "--me", "hex", "--subme", "5", "--bframes", "1", "-o", out, src, "%dx%d" % (w, h)],
                           capture_output=True, text=True, timeout=300, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr[-1000:]
        outs.append(open(out, "rb").read())
    assert len(outs[0]) > 500 and outs[0] == outs[1]


def test_shim_refuses_to_run_without_a_device(pkg, tmp_path):
    """no CPU fallback: with the hooks on and no CUDA device the encoder must fail at open, not fall back to the C tables"""
    if not os.path.exists(CUD):
        pytest.skip("oracle/_ref CLI builds not present")
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a CUDA device is present")
    except ImportError:
        pass
    src = str(tmp_path / "in.yuv")
    open(src, "wb").write(bytes(96 * 64 * 3 // 2))
    r = subprocess.run([CUD, "--qp", "26", "--threads", "1", "-o", str(tmp_path / "o.264"), src, "96x64"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and "no CUDA device" in r.stderr
