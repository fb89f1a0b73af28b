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

# frame sizes have an even macroblock width: for odd widths the reference's lowres planes contain never-written columns (frame.c:301)
CONFIGS = [
    ("config1_dia", 96, 64, 3, "--me dia --subme 1"),
    ("hex_8x8dct_b", 96, 64, 4, "--me hex --subme 5 --8x8dct --bframes 1"),
    ("esa", 96, 64, 2, "--me esa --merange 8 --subme 2"),                                                 # config 2's search
    ("esa_refs_b", 96, 64, 5, "--me esa --merange 16 --subme 2 --ref 2 --bframes 1"),                     # config 2 with two refs and B-frames
    ("cif_esa", 352, 288, 2, "--me esa --merange 16 --subme 2"),                                          # 396 macroblocks per batched launch
    ("esa_p4x4", 64, 48, 2, "--me esa --merange 8 --subme 2 --partitions all"),                           # 4x4 integral plane
    ("umh_rd_weightb", 64, 48, 5, "--me umh --subme 7 --8x8dct --bframes 2 --b-adapt 2 --weightb --mixed-refs --ref 2"),  # config 3
    ("tesa_b3", 64, 48, 5, "--me tesa --merange 8 --subme 6 --bframes 3 --b-adapt 2"),                    # config 4's options
    ("cavlc_8x8dct_deblock", 96, 64, 3, "--me hex --subme 4 --no-cabac --8x8dct --deblock 2:-1 --chroma-qp-offset 3"),
    ("cqm_jvt", 160, 128, 2, "--me hex --subme 4 --cqm jvt"),
    ("crf_aq_lookahead", 96, 64, 5, "--crf 24 --me hex --subme 6 --bframes 2 --b-adapt 2"),               # lowres costs steer rate control
    ("crf_fixed_b", 96, 64, 7, "--crf 24 --me hex --subme 5 --bframes 1 --b-adapt 0 --weightb"),           # B-frame lowres costs, bidir
    ("static_skips", 96, 64, 5, "--me hex --subme 4 --bframes 1 --static"),                               # P- and B-skip probes everywhere
    ("smooth_i16x16", 96, 64, 3, "--me dia --subme 2 --keyint 1 --smooth"),                               # all-intra, Intra16x16 wins often
]


def _clip(pkg, w, h, n, path, static=False, smooth=False):
    from x264_vs2008_b200 import synth
    clip = synth.Clip(w, h, seed=3)
    rng = np.random.default_rng(2)
    with open(path, "wb") as f:
        if smooth:   # smooth surfaces with a few steps: Intra16x16 (V / H / DC / plane) wins on a third to two thirds of the macroblocks
            for i in range(n):
                for sh in (0, 1, 1):
                    yy, xx = np.mgrid[0:h >> sh, 0:w >> sh].astype(np.float64) * (1 << sh)
                    p = 128 + 50 * np.sin(xx / 25 + i) + 40 * np.cos(yy / 17) + np.where((xx // 32 + yy // 32) % 2 == 0, 0, 25)
                    f.write(np.clip(np.round(p), 0, 255).astype(np.uint8).tobytes())
            return
        for i in range(n):
            for p in clip.yuv420(0 if static else i):
                p = np.ascontiguousarray(p).copy()
                if static and i:   # the same picture with a few touched pixels: most macroblocks pass the skip probe, some just fail it
                    for _ in range(p.size // 200):
                        y, x = int(rng.integers(0, p.shape[0])), int(rng.integers(0, p.shape[1]))
                        p[y, x] = np.clip(int(p[y, x]) + int(rng.integers(-12, 13)), 0, 255)
                f.write(p.tobytes())


@pytest.mark.parametrize("tag,w,h,n,opts", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_bitstream_identical(pkg, ctx, tmp_path, tag, w, h, n, opts):
    if not (os.path.exists(REF) and os.path.exists(CUD)):
        pytest.skip("oracle/_ref CLI builds not present (they are produced where the reference sources exist)")
    src = str(tmp_path / "in.yuv")
    static, smooth = "--static" in opts, "--smooth" in opts
    opts = opts.replace(" --static", "").replace(" --smooth", "")
    _clip(pkg, w, h, n, src, static, smooth)
    outs, launches, frames = [], 0, None
    for exe in (REF, CUD):
        out = str(tmp_path / (os.path.basename(exe) + ".264"))
        r = subprocess.run([exe, "--no-asm", "--threads", "1"] + ([] if "--crf" in opts else ["--qp", "26"]) + opts.split() + ["-o", out, src, "%dx%d" % (w, h)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert "encoded %d frames" % n in r.stderr + r.stdout
        outs.append(open(out, "rb").read())
        m = re.search(r"ref_cuda_shim: (\d+) device launches", r.stderr)
        if m:
            launches = int(m.group(1))
        m = re.search(r"frame hooks: (\d+) lowres, (\d+) filter, (\d+) deblock frames recomputed on the device, (\d+) bytes compared equal", r.stderr)
        if m:
            frames = tuple(int(x) for x in m.groups())
    m = re.search(r"me hooks: (\d+) searches, (\d+) qpel refinements, (\d+) bidir refinements repeated on the device and equal; (\d+) left to C", r.stderr)
    me = tuple(int(x) for x in m.groups()) if m else None
    m = re.search(r"residual hooks: (\d+) inter macroblock encodes, (\d+) skip probes repeated on the device and equal; (\d+) encodes left to C", r.stderr)
    resid = tuple(int(x) for x in m.groups()) if m else None
    m = re.search(r"metric hooks: (\d+) AQ frames, (\d+) SSD slabs, (\d+) SSIM slabs repeated on the device and equal", r.stderr)
    metric = tuple(int(x) for x in m.groups()) if m else None
    print(tag, "launches", launches, "frame hooks", frames, "me hooks", me, "residual hooks", resid, "metric hooks", metric)
    m = re.search(r"batched ESA: (\d+) partition searches in (\d+) x264_cuda_me_search_mb launches equal to the C results; (\d+) not representable", r.stderr)
    batched = tuple(int(x) for x in m.groups()) if m else None
    print(tag, "batched ESA", batched)
    m = re.search(r"grid replay: (\d+) ESA searches replayed on device SAD grids and equal; (\d+) needed a vector outside the grid", r.stderr)
    grid = tuple(int(x) for x in m.groups()) if m else None
    print(tag, "grid replay", grid)
    if "--me esa" in opts:   # SAD grids (x264_cuda_sad_grid) + host replay with the exact predictors reproduce each full-pel ESA result
        assert grid is not None and grid[0] > 20 * (n - 1), grid
    if "--me esa" in opts:   # the macroblock-batched kernel searched whole frames of recorded partitions and agreed with every C result
        assert batched is not None and batched[0] > 20 * (n - 1) and batched[1] >= n - 1, batched
    m = re.search(r"intra hooks: (\d+) Intra16x16 decisions, (\d+) chroma mode decisions repeated on the device and equal", r.stderr)
    intra = tuple(int(x) for x in m.groups()) if m else None
    print(tag, "intra hooks", intra)
    if not re.search(r"--subme [6-9]", opts):   # (RD decides modes differently: left to C there) every intra macroblock's 16x16 / chroma mode choice agreed (exit 8)
        assert intra is not None and intra[1] >= (w // 16) * (h // 16) // 2 and (not smooth or intra[0] > n * 2), intra
    m = re.search(r"mc hooks: (\d+) inter macroblocks \((\d+) partition rectangles\) predicted on the device and equal", r.stderr)
    mc = tuple(int(x) for x in m.groups()) if m else None
    print(tag, "mc hooks", mc)
    # the prediction of every inter macroblock encode formed again with x264_cuda_mc_blocks / _bi from the cache's vectors (exit 9)
    assert smooth or static or (mc is not None and mc[0] == resid[0] and mc[1] >= mc[0]), (mc, resid)
    m = re.search(r"lookahead hooks: (\d+) P and (\d+) B frame costs re-evaluated on the device and equal; (\d+) left to C", r.stderr)
    la = tuple(int(x) for x in m.groups()) if m else None
    print(tag, "lookahead hooks", la)
    if "--bframes" in opts and "--b-adapt 0" not in opts:   # B-adapt analysis: every P-type cost estimate it caches, re-evaluated on the device
        assert la is not None and la[0] >= 2, la   # (whether B-type estimates get evaluated depends on the content)
    if "--crf" in opts:   # rate control asks for each frame's lowres cost (x264_rc_analyse_slice): re-evaluated from scratch on the device (exit 10)
        assert la is not None and la[0] >= 1, la
    # PSNR / SSIM slabs of every kept frame, and with rate control the AQ offsets of every input frame (exit 7 on a difference)
    assert metric is not None and metric[1] >= 3 and metric[2] >= 1 and (metric[0] == n or "--crf" not in opts), metric
    # every inter macroblock encode (coefficients, nnz, cbp, reconstruction) and every skip probe was repeated on the device (exit 6 on a difference)
    assert resid is not None and (smooth or (resid[1] > 40 if static else resid[0] >= 4 * (n - 1))), resid
    # every full-resolution motion search of the encode was repeated on the device with the encoder's own predictors and agreed (exit 5 otherwise)
    assert smooth or (me is not None and (static or me[0] > 10 * (n - 1))), me
    assert launches > 1000 * n, launches   # the table entries really ran on the device
    # frame-level hooks: every input frame's lowres planes, every kept reference's deblocking + half-pel/integral planes were recomputed
    # on the device from the encoder's own data, compared byte for byte inside the shim (it exits 4 on a mismatch) and used from then on
    # (lowres planes exist only when the lookahead needs them: B-frame decision or CRF, encoder.c:711-716)
    want_lowres = n if ("--bframes" in opts or "--crf" in opts) else 0
    assert frames is not None and frames[0] == want_lowres and (smooth or (frames[1] >= 1 and frames[2] >= 1 and frames[3] > 0)), frames
    assert len(outs[0]) > 500 and outs[0] == outs[1], (tag, len(outs[0]), len(outs[1]))
