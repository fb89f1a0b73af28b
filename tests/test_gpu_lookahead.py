"""gpu: lowres lookahead (x264_slicetype_frame_cost, S/encoder/slicetype.c:43-355) — intra kernel + wavefront cost kernel
against the oracle over the whole evaluation schedule (I, P dist 1, P dist 2, B with/without cached vectors)."""
import numpy as np
import pytest
import xo_api as X
from helpers import LOOKAHEAD_SCHEDULE, lowres_planes, oracle_lookahead, lookahead_digest

pytestmark = pytest.mark.gpu


def device_lookahead(pkg, ctx, g, clip, me_method, me_range, satd, weighted, n_frames=3, vbv=False, inv_qscale=None):
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
        rows, aq = None, None
        if vbv or inv_qscale is not None:
            score, imbs, isum, aq, rows = ctx.lowres_frame_cost_rc(frames[b], frames[p0], frames[p1], p0, p1, b, vbv=vbv,
                                                                   inv_qscale=inv_qscale[fe] if inv_qscale is not None else None,
                                                                   me_method=me_method, me_range=me_range, flags=flags, do_search=ds,
                                                                   b_intra_calculated=bic)
        else:
            score, imbs, isum = ctx.lowres_frame_cost(frames[b], frames[p0], frames[p1], p0, p1, b, me_method=me_method, me_range=me_range,
                                                      flags=flags, do_search=ds, b_intra_calculated=bic)
        if b < p1:
            score = score * 100 // 120
        d0, d1 = max(b - p0 - 1, 0), max(p1 - b - 1, 0)
        m0, c0, intra = frames[fe].lookahead_get(0, d0)
        m1, c1, _ = frames[fe].lookahead_get(1, d1)
        res.append((name, score, imbs if b == p1 else 0, isum if (b == p1 and p0 != p1) else 0, m0, c0, m1, c1, intra, rows, aq))
    for f in frames:
        f.close()
    return res


@pytest.mark.parametrize("size,aq,vbv", [((176, 144), True, True), ((208, 112), False, True), ((32, 64), True, True), ((176, 144), True, False),
                                         ((1920, 1080), True, True)])
def test_lowres_frame_cost_rc(pkg, ctx, port, size, aq, vbv):
    """the rate-control forms (slicetype.c:300-330): all blocks + per-row sums (VBV), AQ-weighted score"""
    from x264_vs2008_b200 import synth
    w, h = size
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=33)
    planes = lowres_planes(port, g, clip, 3)
    rng = np.random.default_rng(5)
    inv = [rng.integers(128, 512, g.mb_width * g.mb_height).astype(np.uint16) for _ in range(3)] if aq else None
    want = oracle_lookahead(port, g, planes, X.ME_HEX, 16, 1, 0, vbv=vbv, inv_qscale=inv)
    got = device_lookahead(pkg, ctx, g, clip, X.ME_HEX, 16, 1, 0, vbv=vbv, inv_qscale=inv)
    sa, xa = lookahead_digest(want, g, all_blocks=vbv)
    sb, xb = lookahead_digest(got, g, all_blocks=vbv)
    small = g.mb_width <= 2 or g.mb_height <= 2
    for i, (name, *_r) in enumerate(LOOKAHEAD_SCHEDULE):
        bad = np.nonzero(xa[i] != xb[i])[0]
        assert len(bad) == 0, (name, len(bad), bad[:8], xa[i][bad[:8]], xb[i][bad[:8]])
        assert np.array_equal(sa[i], sb[i]), (name, sa[i], sb[i])
        assert want[i][10] == got[i][10], (name, "score_aq", want[i][10], got[i][10])
        if vbv and not small:
            assert np.array_equal(want[i][9], got[i][9]), (name, "row_satd")


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


@pytest.mark.parametrize("size", [(176, 144), (1920, 1080)])
def test_lowres_frame_cost_batch(pkg, ctx, port, size):
    """the P costs cost(i-1, i, i) of a 5-frame window in ONE launch (independent evaluations, interleaved wavefronts) vs the
    oracle run one evaluation at a time; then the B costs between frames 0 and 2 / 2 and 4 in one launch"""
    from x264_vs2008_b200 import synth
    w, h = size
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=52)
    n_frames = 5
    planes = lowres_planes(port, g, clip, n_frames)
    frames = []
    for i in range(n_frames):
        f = ctx.frame(w, h, pkg.FRAME_LOWRES)
        f.upload(clip.luma(i)); f.expand_border(); f.init_lowres(); f.lookahead_alloc(3)
        frames.append(f)
    n = g.mb_width * g.mb_height
    st = [{"mvs": np.zeros((2, 3, n, 2), np.int16), "costs": np.zeros((2, 3, n), np.int32), "intra": np.zeros(n, np.uint16)} for _ in range(n_frames)]

    def oracle_eval(fe, p0, p1, b, ds, bic):
        d0, d1 = max(b - p0 - 1, 0), max(p1 - b - 1, 0)
        s = st[fe]
        state = {"mvs0": s["mvs"][0, d0], "costs0": s["costs"][0, d0], "mvs1": s["mvs"][1, d1], "costs1": s["costs"][1, d1],
                 "intra": s["intra"], "ref1_mvs": st[p1]["mvs"][0, max(p1 - p0 - 1, 0)].copy()}
        o = port.lowres_frame_cost(g, planes[b], planes[p0], planes[p1], p0, p1, b, state, do_search=ds, b_intra_calculated=bic)
        return (o.score, o.intra_mbs if b == p1 else 0, o.intra_cost_sum if b == p1 else 0)

    m = np.zeros((g.mb_height, g.mb_width), bool)
    m[1:-1, 1:-1] = True
    m = m.ravel()
    # batch 1: P costs of frames 1..4 against their predecessors, plus P(0->2) and P(2->4) (distance 2) — all independent
    # batch 0: intra costs of every frame (I evaluations), so that later batches may say b_intra_calculated
    evals = [(i, i, i, i, (0, 0), 0) for i in range(n_frames)]
    want = [oracle_eval(*e) for e in evals]
    got = ctx.lowres_frame_cost_batch([(frames[fe], frames[p0], frames[p1], p0, p1, b, ds, bic) for fe, p0, p1, b, ds, bic in evals])
    assert [a[0] for a in want] == [b[0] for b in got], (want, got)
    evals = [(i, i - 1, i, i, (1, 0), 1) for i in range(1, n_frames)]
    want = [oracle_eval(*e) for e in evals]
    got = ctx.lowres_frame_cost_batch([(frames[fe], frames[p0], frames[p1], p0, p1, b, ds, bic) for fe, p0, p1, b, ds, bic in evals])
    assert want == got, (want, got)
    for fe, p0, p1, b, ds, bic in evals:
        d0 = b - p0 - 1
        mv, cost, _ = frames[fe].lookahead_get(0, d0)
        assert np.array_equal(mv[m], st[fe]["mvs"][0, d0][m]) and np.array_equal(cost[m], st[fe]["costs"][0, d0][m]), (fe, p0)
    # batch 2: distance-2 P costs (their own state slot, dist index 1)
    evals = [(2, 0, 2, 2, (1, 0), 1), (4, 2, 4, 4, (1, 0), 1)]
    want = [oracle_eval(*e) for e in evals]
    got = ctx.lowres_frame_cost_batch([(frames[fe], frames[p0], frames[p1], p0, p1, b, ds, bic) for fe, p0, p1, b, ds, bic in evals])
    assert want == got, (want, got)
    # batch 3: B costs (0,2,1) and (2,4,3): list-0 vectors of frames 1 and 3 are cached from batch 1, list 1 searched now,
    # ref1's list-0 distance-2 vectors come from batch 2
    evals = [(1, 0, 2, 1, (0, 1), 1), (3, 2, 4, 3, (0, 1), 1)]
    want = [oracle_eval(*e) for e in evals]
    got = ctx.lowres_frame_cost_batch([(frames[fe], frames[p0], frames[p1], p0, p1, b, ds, bic) for fe, p0, p1, b, ds, bic in evals])
    assert [a[0] for a in want] == [b[0] for b in got], (want, got)
    for f in frames:
        f.close()


def test_lowres_batch_same_frame_two_distances(pkg, ctx, port):
    """cost(i-1, i, i) and cost(i-2, i, i) search the SAME frame and list at different distances.  The contract allows
    them in one batch (different (frame, list, distance) states); their hand-over words must not collide (one set per distance), or the
    persistent warps spin forever.  Also: a batch that does name the same state twice is refused instead of hanging."""
    from x264_vs2008_b200 import synth
    w, h = 352, 288
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=53)
    n_frames = 3
    planes = lowres_planes(port, g, clip, n_frames)
    frames = []
    for i in range(n_frames):
        f = ctx.frame(w, h, pkg.FRAME_LOWRES)
        f.upload(clip.luma(i)); f.expand_border(); f.init_lowres(); f.lookahead_alloc(3)
        frames.append(f)
    n = g.mb_width * g.mb_height
    st = [{"mvs": np.zeros((2, 3, n, 2), np.int16), "costs": np.zeros((2, 3, n), np.int32), "intra": np.zeros(n, np.uint16)} for _ in range(n_frames)]

    def oracle_eval(fe, p0, p1, b, ds, bic):
        d0, d1 = max(b - p0 - 1, 0), max(p1 - b - 1, 0)
        s = st[fe]
        state = {"mvs0": s["mvs"][0, d0], "costs0": s["costs"][0, d0], "mvs1": s["mvs"][1, d1], "costs1": s["costs"][1, d1],
                 "intra": s["intra"], "ref1_mvs": st[p1]["mvs"][0, max(p1 - p0 - 1, 0)].copy()}
        o = port.lowres_frame_cost(g, planes[b], planes[p0], planes[p1], p0, p1, b, state, do_search=ds, b_intra_calculated=bic)
        return (o.score, o.intra_mbs if b == p1 else 0, o.intra_cost_sum if b == p1 else 0)

    evals = [(i, i, i, i, (0, 0), 0) for i in range(n_frames)]
    for e in evals:
        oracle_eval(*e)
    ctx.lowres_frame_cost_batch([(frames[fe], frames[p0], frames[p1], p0, p1, b, ds, bic) for fe, p0, p1, b, ds, bic in evals])
    evals = [(2, 1, 2, 2, (1, 0), 1), (2, 0, 2, 2, (1, 0), 1), (1, 0, 1, 1, (1, 0), 1)]
    want = [oracle_eval(*e) for e in evals]
    got = ctx.lowres_frame_cost_batch([(frames[fe], frames[p0], frames[p1], p0, p1, b, ds, bic) for fe, p0, p1, b, ds, bic in evals])
    assert want == got, (want, got)
    m = np.zeros((g.mb_height, g.mb_width), bool)
    m[1:-1, 1:-1] = True
    m = m.ravel()
    for d0 in (0, 1):
        mv, cost, _ = frames[2].lookahead_get(0, d0)
        assert np.array_equal(mv[m], st[2]["mvs"][0, d0][m]) and np.array_equal(cost[m], st[2]["costs"][0, d0][m]), d0
    dup = [(2, 1, 2, 2, (1, 0), 1), (2, 1, 2, 2, (1, 0), 1)]
    with pytest.raises(Exception):
        ctx.lowres_frame_cost_batch([(frames[fe], frames[p0], frames[p1], p0, p1, b, ds, bic) for fe, p0, p1, b, ds, bic in dup])
    for f in frames:
        f.close()


def test_lowres_batch_rc_and_refusals(pkg, ctx, port):
    """x264_cuda_lowres_frame_cost_batch_rc: two VBV evaluations (P costs of two frames) in ONE launch against the oracle, and the calls the
    contract refuses: the VBV flag without a row array, VBV and default evaluations in one batch"""
    from x264_vs2008_b200 import synth
    w, h = 352, 288
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=57)
    planes = lowres_planes(port, g, clip, 3)
    frames = []
    for i in range(3):
        f = ctx.frame(w, h, pkg.FRAME_LOWRES)
        f.upload(clip.luma(i)); f.expand_border(); f.init_lowres(); f.lookahead_alloc(3)
        frames.append(f)
    n = g.mb_width * g.mb_height
    rng = np.random.default_rng(9)
    inv = [rng.integers(100, 600, n).astype(np.uint16) for _ in range(3)]
    want = []
    for b in (1, 2):  # cost(b-1, b, b), every block, AQ-weighted
        state = X.lowres_state(g)
        rows = np.zeros(g.mb_height, np.int32)
        o = port.lowres_frame_cost(g, planes[b], planes[b - 1], planes[b], b - 1, b, b, state, do_search=(1, 0), vbv=True, inv_qscale=inv[b], row_satd=rows)
        want.append((o.score, o.intra_mbs, o.intra_cost_sum, o.score_aq, rows, state["mvs0"].copy()))
    got = ctx.lowres_frame_cost_batch_rc([(frames[b], frames[b - 1], frames[b], b - 1, b, b, (1, 0), 0) for b in (1, 2)], inv_qscales=[inv[1], inv[2]])
    for k, b in enumerate((1, 2)):
        assert got[k][:4] == want[k][:4], (b, got[k][:4], want[k][:4])
        assert np.array_equal(got[k][4], want[k][4]), b
        mv, _, _ = frames[b].lookahead_get(0, 0)
        assert np.array_equal(mv, want[k][5]), b  # every block, the frame edge included
    # refusals
    pm = np.array([0, 1, 1, 1, 16, pkg.ME_MBCMP_SATD | pkg.LOWRES_VBV, 1, 0, 1], np.int32)
    res = np.zeros(4, np.int32)
    assert pkg.lib().x264_cuda_lowres_frame_cost_rc(ctx.h, frames[1].h, frames[0].h, frames[1].h, pm.ctypes.data, None, res.ctypes.data, None) == -1
    assert "row_satd" in ctx.error()
    import ctypes as C
    ptrs = [(C.c_void_p * 2)(frames[1].h, frames[2].h), (C.c_void_p * 2)(frames[0].h, frames[1].h), (C.c_void_p * 2)(frames[1].h, frames[2].h)]
    pm2 = np.array([[0, 1, 1, 1, 16, pkg.ME_MBCMP_SATD | pkg.LOWRES_VBV, 1, 0, 1], [1, 2, 2, 1, 16, pkg.ME_MBCMP_SATD, 1, 0, 1]], np.int32)
    res2 = np.zeros((2, 4), np.int32)
    rows = [np.zeros(g.mb_height, np.int32) for _ in range(2)]
    row_p = (C.c_void_p * 2)(rows[0].ctypes.data, rows[1].ctypes.data)
    assert pkg.lib().x264_cuda_lowres_frame_cost_batch_rc(ctx.h, 2, ptrs[0], ptrs[1], ptrs[2], pm2.ctypes.data, None, res2.ctypes.data, row_p) == -1
    assert "cannot share a batch" in ctx.error()
    for f in frames:
        f.close()
