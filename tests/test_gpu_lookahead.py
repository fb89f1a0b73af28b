"""gpu: lowres lookahead (x264_slicetype_frame_cost, S/encoder/slicetype.c:43-355) — intra kernel + wavefront cost kernel
against the oracle over the whole evaluation schedule (I, P dist 1, P dist 2, B with/without cached vectors)."""
import numpy as np
import pytest
import xo_api as X
from helpers import LOOKAHEAD_SCHEDULE, lowres_planes, oracle_lookahead, lookahead_digest

pytestmark = pytest.mark.gpu


def device_lookahead(pkg, ctx, g, clip, me_method, me_range, satd, weighted, n_frames=3):
    frames = []
    for i in range(n_frames):
        f = ctx.frame(g.width, g.height, pkg.FRAME_LOWRES)
        f.upload(clip.luma(i))
        f.expand_border()
        f.init_lowres()
        f.lookahead_alloc(3)
        frames.append(f)
    flags = (pkg.ME_MBCMP_SATD if satd else 0) | (pkg.LOWRES_WEIGHTED_BIPRED if weighted else 0)
    res = []
    for name, fe, p0, p1, b, ds, bic in LOOKAHEAD_SCHEDULE:
        score, imbs, isum = ctx.lowres_frame_cost(frames[b], frames[p0], frames[p1], p0, p1, b, me_method=me_method, me_range=me_range,
                                                  flags=flags, do_search=ds, b_intra_calculated=bic)
        if b < p1:
            score = score * 100 // 120
        d0, d1 = max(b - p0 - 1, 0), max(p1 - b - 1, 0)
        m0, c0, intra = frames[fe].lookahead_get(0, d0)
        m1, c1, _ = frames[fe].lookahead_get(1, d1)
        res.append((name, score, imbs if b == p1 else 0, isum if (b == p1 and p0 != p1) else 0, m0, c0, m1, c1, intra))
    for f in frames:
        f.close()
    return res


@pytest.mark.parametrize("size,method,satd,weighted", [((176, 144), X.ME_HEX, 1, 0), ((176, 144), X.ME_DIA, 0, 0), ((208, 112), X.ME_HEX, 1, 1),
                                                       ((32, 64), X.ME_HEX, 1, 0), ((640, 352), X.ME_HEX, 1, 1), ((1920, 1080), X.ME_HEX, 1, 0)])
def test_lowres_frame_cost(pkg, ctx, port, size, method, satd, weighted):
    from x264_vs2008_b200 import synth
    w, h = size
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=31)
    planes = lowres_planes(port, g, clip, 3)
    want = oracle_lookahead(port, g, planes, method, 16, satd, weighted)
    got = device_lookahead(pkg, ctx, g, clip, method, 16, satd, weighted)
    sa, xa = lookahead_digest(want, g)
    sb, xb = lookahead_digest(got, g)
    for i, (name, *_r) in enumerate(LOOKAHEAD_SCHEDULE):
        bad = np.nonzero(xa[i] != xb[i])[0]
        assert len(bad) == 0, (name, len(bad), bad[:8], xa[i][bad[:8]], xb[i][bad[:8]])
        assert np.array_equal(sa[i], sb[i]), (name, sa[i], sb[i])


def test_lowres_intra_all_blocks(pkg, ctx, port):
    """the intra kernel evaluates every block (also the frame-edge ones the reference skips): check all of them vs the oracle"""
    from x264_vs2008_b200 import synth
    w, h = 352, 288
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=77)
    planes = lowres_planes(port, g, clip, 1)
    for satd in (0, 1):
        f = ctx.frame(w, h, pkg.FRAME_LOWRES)
        f.upload(clip.luma(0)); f.expand_border(); f.init_lowres(); f.lookahead_alloc(1)
        ctx.lowres_frame_cost(f, f, f, 0, 0, 0, flags=pkg.ME_MBCMP_SATD if satd else 0, do_search=(0, 0))
        _, _, intra = f.lookahead_get(0, 0)
        f.close()
        want = np.array([port.lib.xo_lowres_intra_cost(X._ptr(planes[0][0], X.u8p, g.origin_lowres), g.stride_lowres, 8 * (i % g.mb_width), 8 * (i // g.mb_width), satd)
                         for i in range(g.mb_width * g.mb_height)])
        assert np.array_equal(intra, want), np.nonzero(intra != want)[0][:10]
