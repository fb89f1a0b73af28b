"""gpu: Intra16x16 + chroma 8x8 candidate costs from neighbouring macroblocks vs the oracle, bit-exact"""
import numpy as np
import pytest
import xo_api as X

pytestmark = pytest.mark.gpu


def _nb(pad, p, x0, y0, n):
    """neighbour vector (corner, row above, left column) of the n x n block at (x0, y0) of a plane padded by p"""
    return np.concatenate([pad[p + y0 - 1, p + x0 - 1:p + x0 - 1 + 1], pad[p + y0 - 1, p + x0:p + x0 + n], pad[p + y0:p + y0 + n, p + x0 - 1]]).astype(np.uint8)


def _check(pkg, ctx, port, fenc_yuv, fdec_yuv, jobs, tag):
    (ey, eu, ev), (dy, du, dv) = fenc_yuv, fdec_yuv
    h, w = ey.shape
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    fenc.upload(ey); fenc.upload_chroma(eu, ev)
    fdec.upload(dy); fdec.upload_chroma(du, dv); fdec.expand_border()
    got = ctx.intra_mb_costs(fenc, fdec, jobs)
    py, pu, pv = np.pad(dy, 1, mode="edge"), np.pad(du, 1, mode="edge"), np.pad(dv, 1, mode="edge")
    wins16, winsc = set(), set()
    for i, j in enumerate(jobs):
        mx, my = int(j["mb_x"]), int(j["mb_y"])
        tile = lambda a, n: np.ascontiguousarray(a[my * n:my * n + n, mx * n:mx * n + n])
        iin = X.IntraIn(int(j["neighbour"]), int(j["lambda"]), int(j["flags"]) & 1, (int(j["flags"]) >> 1) & 1)
        o = port.intra_mb_costs(iin, tile(ey, 16), tile(eu, 8), tile(ev, 8), _nb(py, 1, mx * 16, my * 16, 16), _nb(pu, 1, mx * 8, my * 8, 8),
                                _nb(pv, 1, mx * 8, my * 8, 8))
        g = got[i]
        mine = (tuple(int(v) for v in g["cost16"]), tuple(int(v) for v in g["cost_chroma"]), int(g["best16"]), int(g["best_chroma"]), int(g["mode16"]),
                int(g["mode_chroma"]))
        assert mine == o.astuple(), (tag, i, mx, my, int(j["neighbour"]), int(j["lambda"]), int(j["flags"]))
        wins16.add(o.mode16); winsc.add(o.mode_chroma)
    fenc.close(); fdec.close()
    return wins16, winsc


def test_intra_mb_costs_frame(pkg, ctx, port):
    """every macroblock of a CIF frame with the neighbour mask its position implies (the I-slice case), SATD and SAD, P and B lambdas"""
    from x264_vs2008_b200 import synth
    w, h = 352, 288
    mbw, mbh = w // 16, h // 16
    clip = synth.Clip(w, h, seed=23)
    ey, eu, ev = clip.yuv420(1)
    rng = np.random.default_rng(5)
    noisy = lambda a: np.clip(a.astype(np.int32) + rng.integers(-3, 4, a.shape), 0, 255).astype(np.uint8)
    fdec = (noisy(ey), noisy(eu), noisy(ev))  # a reconstruction-like picture
    for rep, flags in enumerate((pkg.INTRA_SATD, 0, pkg.INTRA_SATD | pkg.INTRA_SLICE_B)):
        jobs = np.zeros(mbw * mbh, pkg.INTRA_JOB)
        for i in range(len(jobs)):
            mx, my = i % mbw, i // mbw
            nbr = (pkg.MB_LEFT if mx else 0) | (pkg.MB_TOP if my else 0) | (pkg.MB_TOPLEFT if mx and my else 0) | (pkg.MB_TOPRIGHT if my and mx < mbw - 1 else 0)
            jobs[i] = (mx, my, nbr, flags, pkg.lib().x264_cuda_host_lambda(int(rng.integers(10, 52))))
        w16, wc = _check(pkg, ctx, port, (ey, eu, ev), fdec, jobs, rep)
        assert len(w16) >= 5 and len(wc) >= 5, (w16, wc)


def test_intra_mb_costs_synthetic(pkg, ctx, port):
    """ramps / flats / extremes laid out as a frame, arbitrary neighbour masks (slice boundaries, constrained intra): every mode of both
    enums wins somewhere"""
    w, h = 320, 240
    mbw, mbh = w // 16, h // 16
    rng = np.random.default_rng(6)
    yy, xx = np.mgrid[0:h, 0:w]
    ey = np.zeros((h, w), np.int32)
    for my in range(mbh):
        for mx in range(mbw):
            k = (mx + my * mbw) % 5
            sl = np.s_[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16]
            lx, ly = xx[sl] - mx * 16, yy[sl] - my * 16
            if k == 0:
                ey[sl] = rng.integers(0, 256, (16, 16))
            elif k == 1:
                ey[sl] = rng.integers(0, 256) + rng.integers(-20, 21) * lx + rng.integers(-20, 21) * ly
            elif k == 2:
                ey[sl] = rng.integers(0, 256) + rng.integers(-2, 3, (16, 16))
            elif k == 3:
                ey[sl] = rng.integers(0, 2, (16, 16)) * 255
            else:
                ey[sl] = 128 + rng.integers(-4, 5) * lx + rng.integers(-4, 5) * ly
    ey = np.clip(ey, 0, 255).astype(np.uint8)
    eu = np.ascontiguousarray(ey[::2, ::2]); ev = np.ascontiguousarray(255 - ey[1::2, 1::2])
    # the "reconstruction": the same picture shifted by one pixel diagonally, so neighbours continue each block's ramp
    dy = np.roll(ey, (1, 1), (0, 1)); du = np.roll(eu, (1, 1), (0, 1)); dv = np.roll(ev, (1, 1), (0, 1))
    jobs = np.zeros(mbw * mbh, pkg.INTRA_JOB)
    for i in range(len(jobs)):
        jobs[i] = (i % mbw, i // mbw, int(rng.integers(0, 16)) if i % 3 else 0xF, int(rng.integers(0, 4)), int(rng.choice([1, 2, 4, 10, 32, 91])))
    w16, wc = _check(pkg, ctx, port, (ey, eu, ev), (dy, du, dv), jobs, "synthetic")
    assert w16 == set(range(7)) and wc == set(range(7)), (w16, wc)
