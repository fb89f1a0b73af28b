"""gpu: frame-batched motion compensation (luma qpel + chroma 1/8 pel) vs the oracle"""
import numpy as np
import pytest
import xo_api as X

pytestmark = pytest.mark.gpu


def test_mc_blocks(pkg, ctx, port):
    from x264_vs2008_b200 import synth
    w, h = 352, 288
    clip = synth.Clip(w, h, seed=13)
    y0, u0, v0 = clip.yuv420(0)
    g = port.geometry(w, h)
    flags = pkg.FRAME_HPEL | pkg.FRAME_CHROMA
    fref, fdec = ctx.frame(w, h, flags), ctx.frame(w, h, flags)
    fref.upload(y0); fref.expand_border(); fref.filter()
    # chroma with replicated 16-px borders, uploaded including the border (origin at -16,-16 is not addressable through
    # upload_chroma, so pad inside the picture: test vectors stay >= 2 samples inside)
    fref.upload_chroma(u0, v0)
    plane = port.plane_from_picture(g, y0)
    fh, fv, fc, _ = port.frame_filter(g, plane, 0, want_integral=False)
    planes = [plane, fh, fv, fc]
    rng = np.random.default_rng(3)
    n = 2000
    jobs = np.zeros(n, pkg.MC_JOB)
    for i in range(n):
        bw, bh = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)][int(rng.integers(0, 7))]
        bx = int(rng.integers(2, (w - 40) // 4)) * 4 + 8
        by = int(rng.integers(2, (h - 40) // 4)) * 4 + 8
        jobs[i] = (bx, by, int(rng.integers(-30, 31)), int(rng.integers(-30, 31)), bw, bh, (0, 0))
    # blocks overlap in fdec; process one at a time against the oracle by running disjoint batches
    got_y, got_u, got_v = None, None, None
    for i in range(0, n, 1):
        if i % 50:
            continue
        batch = jobs[i:i + 1]
        ctx.mc_blocks(fref, fdec, batch)
        j = batch[0]
        bx, by, mvx, mvy, bw, bh = int(j["bx"]), int(j["by"]), int(j["mvx"]), int(j["mvy"]), int(j["w"]), int(j["h"])
        got_y = fdec.download(pkg.PLANE_FULL)[32 + by:32 + by + bh, 32 + bx:32 + bx + bw]
        want = np.zeros((bh, 16), np.uint8)
        arr = (X.u8p * 4)(*[X._ptr(p, X.u8p, g.origin + by * g.stride + bx) for p in planes])
        port.lib.xo_mc_luma(X._ptr(want), 16, arr, g.stride, mvx, mvy, bw, bh)
        assert np.array_equal(got_y, want[:, :bw]), (i, "luma")
        for pl, src in ((pkg.PLANE_CB, u0), (pkg.PLANE_CR, v0)):
            got_c = fdec.download(pl)[16 + by // 2:16 + by // 2 + bh // 2, 16 + bx // 2:16 + bx // 2 + bw // 2]
            wc = np.zeros((bh // 2, 8), np.uint8)
            srcp = np.ascontiguousarray(src)
            port.lib.xo_mc_chroma(X._ptr(wc), 8, X._ptr(srcp, X.u8p, (by // 2) * srcp.shape[1] + bx // 2), srcp.shape[1], mvx, mvy, bw // 2, bh // 2)
            assert np.array_equal(got_c, wc[:, :bw // 2]), (i, "chroma")
    fref.close(); fdec.close()


def test_mc_blocks_bi(pkg, ctx, port):
    """x264_mb_mc_01xywh: both lists' luma (get_ref) and chroma (mc_chroma) predictions blended by mc.avg — plain average and the
    implicit-weighted-bipred weights, all seven block sizes"""
    from x264_vs2008_b200 import synth
    from helpers import padded_chroma
    w, h = 352, 288
    clip = synth.Clip(w, h, seed=17)
    (y0, u0, v0), (y1, u1, v1) = clip.yuv420(0), clip.yuv420(2)
    g = port.geometry(w, h)
    flags = pkg.FRAME_HPEL | pkg.FRAME_CHROMA
    f0, f1, fdec = ctx.frame(w, h, flags), ctx.frame(w, h, flags), ctx.frame(w, h, flags)
    refs = []
    for f, (yy, uu, vv) in ((f0, (y0, u0, v0)), (f1, (y1, u1, v1))):
        f.upload(yy); f.upload_chroma(uu, vv); f.expand_border(); f.filter()
        plane = port.plane_from_picture(g, yy)
        fh, fv, fc, _ = port.frame_filter(g, plane, 0, want_integral=False)
        refs.append(([plane, fh, fv, fc], padded_chroma(g, uu), padded_chroma(g, vv)))
    rng = np.random.default_rng(4)
    sizes = [(16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4)]
    for t in range(60):
        ip = int(rng.integers(0, 7))
        bw, bh = sizes[ip]
        bx, by = int(rng.integers(0, (w - 16) // 4)) * 4, int(rng.integers(0, (h - 16) // 4)) * 4
        if bw >= 8:
            bx, by = bx & ~7, by & ~7
        weight = int(rng.choice([32, 32, 21, 43, 16, 48, 11, 53]))
        job = np.zeros(1, pkg.MC_BI_JOB)
        job[0]["bx"], job[0]["by"], job[0]["w"], job[0]["h"], job[0]["weight"] = bx, by, bw, bh, weight
        job[0]["mv0"] = [int(rng.integers(-60, 61)), int(rng.integers(-40, 41))]
        job[0]["mv1"] = [int(rng.integers(-60, 61)), int(rng.integers(-40, 41))]
        ctx.mc_blocks_bi(f0, f1, fdec, job)
        tmp = [np.zeros((16, 16), np.uint8), np.zeros((16, 16), np.uint8)]
        for l in range(2):
            arr = (X.u8p * 4)(*[X._ptr(p, X.u8p, g.origin + by * g.stride + bx) for p in refs[l][0]])
            mv = job[0]["mv0"] if l == 0 else job[0]["mv1"]
            port.lib.xo_mc_luma(X._ptr(tmp[l]), 16, arr, g.stride, int(mv[0]), int(mv[1]), bw, bh)
        want = np.zeros((16, 16), np.uint8)
        port.lib.xo_pixel_avg(ip, X._ptr(want), 16, X._ptr(tmp[0]), 16, X._ptr(tmp[1]), 16, weight)
        got = fdec.download(pkg.PLANE_FULL)[32 + by:32 + by + bh, 32 + bx:32 + bx + bw]
        assert np.array_equal(got, want[:bh, :bw]), (t, "luma", ip, weight)
        for k, pl in ((1, pkg.PLANE_CB), (2, pkg.PLANE_CR)):
            for l in range(2):
                cp = refs[l][k]
                mv = job[0]["mv0"] if l == 0 else job[0]["mv1"]
                port.lib.xo_mc_chroma(X._ptr(tmp[l]), 16, X._ptr(cp, X.u8p, (16 + by // 2) * cp.shape[1] + 16 + bx // 2), cp.shape[1], int(mv[0]), int(mv[1]),
                                      bw // 2, bh // 2)
            cw, ch = bw // 2, bh // 2
            ipc = {(8, 8): 3, (8, 4): 4, (4, 8): 5, (4, 4): 6, (4, 2): 7, (2, 4): 8, (2, 2): 9}[(cw, ch)]
            port.lib.xo_pixel_avg(ipc, X._ptr(want), 16, X._ptr(tmp[0]), 16, X._ptr(tmp[1]), 16, weight)
            gotc = fdec.download(pl)[16 + by // 2:16 + by // 2 + ch, 16 + bx // 2:16 + bx // 2 + cw]
            assert np.array_equal(gotc, want[:ch, :cw]), (t, "chroma", k, ip, weight)
    for f in (f0, f1, fdec):
        f.close()
