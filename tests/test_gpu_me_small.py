"""gpu: DIA / HEX searches, '-> qpel mv' and refine_subpel (x264_cuda_me_search_small) and the packed block metrics,
vs the oracle, bit-exact"""
import numpy as np
import pytest
import xo_api as X
from helpers import make_me_jobs

pytestmark = pytest.mark.gpu


def _setup(pkg, ctx, port, w, h, seed):
    from x264_vs2008_b200 import synth
    clip = synth.Clip(w, h, seed=seed)
    g = port.geometry(w, h)
    fenc, fref = ctx.frame(w, h, 0), ctx.frame(w, h, pkg.FRAME_HPEL)
    fenc.upload(clip.luma(1)); fenc.expand_border()
    fref.upload(clip.luma(0)); fref.expand_border(); fref.filter()
    pe, pr = port.plane_from_picture(g, clip.luma(1)), port.plane_from_picture(g, clip.luma(0))
    fh, fv, fc, _ = port.frame_filter(g, pr, 0, want_integral=False)
    return g, fenc, fref, pe, [pr, fh, fv, fc]


def _fill_spel(jobs, mis):
    for j, mi in zip(jobs, mis):
        j["mv_min_spel"] = [mi.mv_min_spel[0], mi.mv_min_spel[1]]
        j["mv_max_spel"] = [mi.mv_max_spel[0], mi.mv_max_spel[1]]


@pytest.mark.parametrize("method", [X.ME_DIA, X.ME_HEX])
@pytest.mark.parametrize("subme", [1, 2, 3, 4, 5, 7])
def test_small_search(pkg, ctx, port, method, subme):
    w, h = 320, 192
    g, fenc, fref, pe, planes = _setup(pkg, ctx, port, w, h, seed=40 + subme)
    for mbcmp_satd in (0, 1):
        jobs, mis = make_me_jobs(pkg, g, seed=100 * method + subme, n=400, me_range=16, qp=(12, 26, 38), pixels=(0, 1, 2, 3, 4, 5, 6),
                                 mvp_spread=40)
        _fill_spel(jobs, mis)
        jobs["flags"] = pkg.ME_MBCMP_SATD if mbcmp_satd else 0
        res = ctx.me_search_small(fenc, fref, method, 16, subme, jobs)
        bad = []
        for i, mi in enumerate(mis):
            mi.me_method = method
            o = port.me_search_subpel(g, pe, planes, None, mi, subme, mbcmp_satd)
            got = (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"]), int(res[i]["bmx"]), int(res[i]["bmy"]))
            want = (o.mv[0], o.mv[1], o.cost, o.cost_mv, o.bmx, o.bmy)
            if got != want:
                bad.append((i, mi.i_pixel, got, want))
        assert not bad, (mbcmp_satd, len(bad), bad[:4])
    fenc.close(); fref.close()


@pytest.mark.parametrize("subme", [1, 2])
def test_esa_then_subpel(pkg, ctx, port, subme):
    """full x264_me_search_ref for --me esa: ESA kernel -> seeded small-search kernel (me.c:603-631)"""
    w, h = 320, 192
    g, fenc, fref, pe, planes = _setup(pkg, ctx, port, w, h, seed=77)
    jobs, mis = make_me_jobs(pkg, g, seed=5, n=400, me_range=16, qp=(20, 30), pixels=(0, 1, 2, 3))
    _fill_spel(jobs, mis)
    jobs["flags"] = pkg.ME_MBCMP_SATD
    fp = ctx.me_search(fenc, fref, 16, jobs)
    j2 = jobs.copy()
    j2["seed_mv"][:, 0], j2["seed_mv"][:, 1], j2["seed_cost"] = fp["bmx"], fp["bmy"], fp["bcost"]
    res = ctx.me_search_small(fenc, fref, pkg.ME_METHOD_SEEDED, 16, subme, j2)
    for i, mi in enumerate(mis):
        o = port.me_search_subpel(g, pe, planes, None, mi, subme, 1)
        assert (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"])) == (o.mv[0], o.mv[1], o.cost, o.cost_mv), i
    fenc.close(); fref.close()


def test_block_cmp(pkg, ctx, port):
    """checkasm-style (checkasm.c:222-295): random tiles + the max-difference 'overflow' patterns"""
    rng = np.random.default_rng(8)
    n = 600
    a = rng.integers(0, 256, (n, 16, 16), dtype=np.uint8)
    b = rng.integers(0, 256, (n, 16, 16), dtype=np.uint8)
    yy, xx = np.mgrid[0:16, 0:16]
    a[0], b[0] = 0, 255
    a[1], b[1] = ((xx + yy) & 1) * 255, ((xx + yy + 1) & 1) * 255
    a[2], b[2] = (xx & 1) * 255, (yy & 1) * 255
    a[3], b[3] = ((xx >> 1) & 1) * 255, ((yy >> 2) & 1) * 255
    a[4] = b[4]
    for metric in range(4):
        for ip in range(7):
            if metric == X.SA8D and ip not in (0, 3):
                continue
            got = ctx.block_cmp(metric, ip, a, b)
            want = [port.pixel_cmp(metric, ip, a[i], 16, b[i], 16) for i in range(n)]
            assert list(got) == want, (metric, ip)


@pytest.mark.parametrize("me_range,subme,fpel_satd", [(16, 1, 0), (16, 2, 1), (16, 7, 1), (8, 1, 0), (24, 2, 1), (32, 1, 0)])
def test_tesa(pkg, ctx, port, me_range, subme, fpel_satd):
    """--me tesa end to end on the device: ADS/SAD thresholds, keeper list, SATD on the keepers, sub-pel tail"""
    from x264_vs2008_b200 import synth
    w, h = 320, 192
    clip = synth.Clip(w, h, seed=61)
    g = port.geometry(w, h)
    fenc = ctx.frame(w, h, 0)
    fref = ctx.frame(w, h, pkg.FRAME_HPEL | pkg.FRAME_INTEGRAL | pkg.FRAME_INTEGRAL4)
    fenc.upload(clip.luma(1)); fenc.expand_border()
    fref.upload(clip.luma(0)); fref.expand_border(); fref.filter()
    pe, pr = port.plane_from_picture(g, clip.luma(1)), port.plane_from_picture(g, clip.luma(0))
    fh, fv, fc, integ = port.frame_filter(g, pr, 1)
    jobs, mis = make_me_jobs(pkg, g, seed=me_range + subme, n=400, me_range=me_range, qp=(12, 26, 38), pixels=(0, 1, 2, 3, 4, 5, 6),
                             tesa=True, fpel_satd=bool(fpel_satd))
    _fill_spel(jobs, mis)
    mbs = 1 if subme > 1 else 0
    jobs["flags"] = (pkg.ME_FPEL_SATD if fpel_satd else 0) | (pkg.ME_MBCMP_SATD if mbs else 0)
    res = ctx.me_search_small(fenc, fref, pkg.ME_METHOD_TESA, me_range, subme, jobs)
    bad = []
    for i, mi in enumerate(mis):
        mi.b_sub8x8 = 1
        o = port.me_search_subpel(g, pe, [pr, fh, fv, fc], integ, mi, subme, mbs)
        got = (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"]), int(res[i]["bmx"]), int(res[i]["bmy"]))
        want = (o.mv[0], o.mv[1], o.cost, o.cost_mv, o.bmx, o.bmy)
        if got != want:
            bad.append((i, mi.i_pixel, got, want))
    assert not bad, (len(bad), bad[:4])
    fenc.close(); fref.close()


@pytest.mark.parametrize("method", [X.ME_DIA, X.ME_HEX])
@pytest.mark.parametrize("subme", [5, 6, 7])
def test_chroma_me(pkg, ctx, port, method, subme):
    """b_chroma_me: COST_MV_SATD adds mc_chroma + mbcmp[i_pixel+3] of U and V (me.c:655-677) for partitions >= 8x8; also exercises
    the device-side chroma border expansion (frame.c:229-236)"""
    from x264_vs2008_b200 import synth
    from helpers import padded_chroma
    w, h = 320, 192
    clip = synth.Clip(w, h, seed=60 + subme)
    g = port.geometry(w, h)
    (y1, u1, v1), (y0, u0, v0) = clip.yuv420(1), clip.yuv420(0)
    fenc, fref = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_HPEL | pkg.FRAME_CHROMA)
    fenc.upload(y1); fenc.upload_chroma(u1, v1); fenc.expand_border()
    fref.upload(y0); fref.upload_chroma(u0, v0); fref.expand_border(); fref.filter()
    chroma = [padded_chroma(g, c) for c in (u1, v1, u0, v0)]
    for k, pl in enumerate((pkg.PLANE_CB, pkg.PLANE_CR)):  # device border expansion of the chroma planes
        assert np.array_equal(fref.download(pl), chroma[2 + k])
    pe, pr = port.plane_from_picture(g, y1), port.plane_from_picture(g, y0)
    fh, fv, fc, _ = port.frame_filter(g, pr, 0, want_integral=False)
    jobs, mis = make_me_jobs(pkg, g, seed=200 * method + subme, n=400, me_range=16, qp=(12, 26, 38), pixels=(0, 1, 2, 3, 4, 5, 6), mvp_spread=40)
    _fill_spel(jobs, mis)
    for j, mi in zip(jobs, mis):  # chroma needs even block positions; partitions >= 8x8 sit on multiples of 8 anyway
        j["bx"], j["by"] = (int(j["bx"]) // 8) * 8, (int(j["by"]) // 8) * 8
        mi.bx, mi.by = int(j["bx"]), int(j["by"])
    jobs["flags"] = pkg.ME_MBCMP_SATD | pkg.ME_CHROMA
    res = ctx.me_search_small(fenc, fref, method, 16, subme, jobs)
    bad, n_differs = [], 0
    for i, mi in enumerate(mis):
        mi.me_method = method
        o = port.me_search_subpel_chroma(g, pe, [pr, fh, fv, fc], None, chroma, mi, subme, 1)
        plain = port.me_search_subpel(g, pe, [pr, fh, fv, fc], None, mi, subme, 1)
        n_differs += (plain.cost != o.cost)
        got = (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"]))
        if got != (o.mv[0], o.mv[1], o.cost, o.cost_mv):
            bad.append((i, mi.i_pixel, got, (o.mv[0], o.mv[1], o.cost, o.cost_mv)))
    assert not bad, (len(bad), bad[:4])
    assert n_differs > 100
    fenc.close(); fref.close()


def test_tesa_many(pkg, ctx, port):
    """4000 TESA searches on a 720p pair (frame-edge macroblocks, clipped windows, all partition sizes): guards the keeper-list /
    threshold logic at scale (a 1-in-400 miscompile of this path was found and fixed in round 1)"""
    from x264_vs2008_b200 import synth
    w, h = 1280, 720
    clip = synth.Clip(w, h, seed=88)
    g = port.geometry(w, h)
    fenc = ctx.frame(w, h, 0)
    fref = ctx.frame(w, h, pkg.FRAME_HPEL | pkg.FRAME_INTEGRAL | pkg.FRAME_INTEGRAL4)
    fenc.upload(clip.luma(1)); fenc.expand_border()
    fref.upload(clip.luma(0)); fref.expand_border(); fref.filter()
    pe, pr = port.plane_from_picture(g, clip.luma(1)), port.plane_from_picture(g, clip.luma(0))
    fh, fv, fc, integ = port.frame_filter(g, pr, 1)
    jobs, mis = make_me_jobs(pkg, g, seed=4242, n=4000, me_range=16, qp=(12, 26, 38), pixels=(0, 1, 2, 3, 4, 5, 6), tesa=True, fpel_satd=True)
    _fill_spel(jobs, mis)
    jobs["flags"] = pkg.ME_FPEL_SATD | pkg.ME_MBCMP_SATD
    res = ctx.me_search_small(fenc, fref, pkg.ME_METHOD_TESA, 16, 7, jobs)
    bad = []
    for i, mi in enumerate(mis):
        mi.b_sub8x8 = 1
        o = port.me_search_subpel(g, pe, [pr, fh, fv, fc], integ, mi, 7, 1)
        got = (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"]), int(res[i]["bmx"]), int(res[i]["bmy"]))
        if got != (o.mv[0], o.mv[1], o.cost, o.cost_mv, o.bmx, o.bmy):
            bad.append((i, mi.i_pixel, got, (o.mv[0], o.mv[1], o.cost, o.cost_mv, o.bmx, o.bmy)))
    assert not bad, (len(bad), bad[:4])
    fenc.close(); fref.close()


@pytest.mark.parametrize("me_range", [16, 24, 8])
@pytest.mark.parametrize("subme", [1, 2, 5])
def test_umh(pkg, ctx, port, me_range, subme):
    """--me umh on the device (me.c:306-447): every early-termination path, the adaptive range contexts, the hexagon grid, hex2"""
    w, h = 320, 192
    g, fenc, fref, pe, planes = _setup(pkg, ctx, port, w, h, seed=90 + subme)
    for spread, centre in ((48, None), (10, (-20, -12)), (4, (-20, -12))):
        jobs, mis = make_me_jobs(pkg, g, seed=700 + me_range + spread, n=400, me_range=me_range, qp=(12, 26, 38), pixels=(0, 1, 2, 3, 4, 5, 6),
                                 mvp_spread=spread, centre=centre)
        _fill_spel(jobs, mis)
        for i, (j, mi) in enumerate(zip(jobs, mis)):
            mi.me_method = X.ME_UMH
            if centre is not None and i % 2:  # neighbours agreeing with the predictor: the small-range contexts
                for k in range(mi.i_mvc):
                    mi.mvc[k][0], mi.mvc[k][1] = mi.mvp[0] + (k % 3) - 1, mi.mvp[1] + (k % 2)
                    j["mvc"][k] = [mi.mvc[k][0], mi.mvc[k][1]]
        jobs["flags"] = pkg.ME_MBCMP_SATD
        res = ctx.me_search_small(fenc, fref, pkg.ME_METHOD_UMH, me_range, subme, jobs)
        bad = []
        for i, mi in enumerate(mis):
            o = port.me_search_subpel(g, pe, planes, None, mi, subme, 1)
            got = (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"]), int(res[i]["bmx"]), int(res[i]["bmy"]))
            want = (o.mv[0], o.mv[1], o.cost, o.cost_mv, o.bmx, o.bmy)
            if got != want:
                bad.append((i, mi.i_pixel, got, want))
        assert not bad, (spread, len(bad), bad[:4])
    fenc.close(); fref.close()


@pytest.mark.parametrize("with_chroma", [0, 1])
def test_refine_qpel(pkg, ctx, port, with_chroma):
    """x264_me_refine_qpel on the device: refine_subpel with b_refine_qpel = 1 from a given quarter-pel vector and cost"""
    from x264_vs2008_b200 import synth
    from helpers import padded_chroma
    w, h = 320, 192
    clip = synth.Clip(w, h, seed=71)
    g = port.geometry(w, h)
    (y1, u1, v1), (y0, u0, v0) = clip.yuv420(1), clip.yuv420(0)
    fenc, fref = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_HPEL | pkg.FRAME_CHROMA)
    fenc.upload(y1); fenc.upload_chroma(u1, v1); fenc.expand_border()
    fref.upload(y0); fref.upload_chroma(u0, v0); fref.expand_border(); fref.filter()
    chroma = [padded_chroma(g, c) for c in (u1, v1, u0, v0)] if with_chroma else None
    pe, pr = port.plane_from_picture(g, y1), port.plane_from_picture(g, y0)
    fh, fv, fc, _ = port.frame_filter(g, pr, 0, want_integral=False)
    rng = np.random.default_rng(9)
    for subme in (1, 2, 3, 5, 7):
        jobs, mis = make_me_jobs(pkg, g, seed=40 + subme, n=300, me_range=16, qp=(12, 26, 38), pixels=(0, 1, 2, 3, 4, 5, 6), mvp_spread=30, centre=(-20, -12))
        _fill_spel(jobs, mis)
        for j, mi in zip(jobs, mis):
            j["bx"], j["by"] = (int(j["bx"]) // 8) * 8 if mi.i_pixel <= 3 else j["bx"], (int(j["by"]) // 8) * 8 if mi.i_pixel <= 3 else j["by"]
            mi.bx, mi.by = int(j["bx"]), int(j["by"])
            j["seed_mv"] = [int(rng.integers(-12, 13)) - 20, int(rng.integers(-12, 13)) - 12]
            j["seed_cost"] = int(rng.integers(200, 6000))
        jobs["flags"] = pkg.ME_MBCMP_SATD | (pkg.ME_CHROMA if with_chroma else 0)
        res = ctx.me_search_small(fenc, fref, pkg.ME_METHOD_REFINE_QPEL, 16, subme, jobs)
        bad = []
        for i, (j, mi) in enumerate(zip(jobs, mis)):
            o = port.me_refine_qpel(g, pe, [pr, fh, fv, fc], chroma, mi, subme, 1, [int(j["seed_mv"][0]), int(j["seed_mv"][1])], int(j["seed_cost"]))
            got = (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"]))
            if got != (o.mv[0], o.mv[1], o.cost, o.cost_mv):
                bad.append((i, mi.i_pixel, got, (o.mv[0], o.mv[1], o.cost, o.cost_mv)))
        assert not bad, (subme, len(bad), bad[:4])
    fenc.close(); fref.close()


def test_refine_bidir(pkg, ctx, port):
    """x264_me_refine_bidir_satd on the device vs the oracle (which is pinned to the reference by tests/test_oracle_vs_ref.py)"""
    from x264_vs2008_b200 import synth
    from test_oracle_vs_ref import bidir_cases
    w, h = 320, 192
    clip = synth.Clip(w, h, seed=81)
    g = port.geometry(w, h)
    fenc = ctx.frame(w, h, 0)
    fenc.upload(clip.luma(1)); fenc.expand_border()
    pe = port.plane_from_picture(g, clip.luma(1))
    frefs, refs = [], []
    for fr in (0, 2):
        f = ctx.frame(w, h, pkg.FRAME_HPEL)
        f.upload(clip.luma(fr)); f.expand_border(); f.filter()
        frefs.append(f)
        pr = port.plane_from_picture(g, clip.luma(fr))
        fh, fv, fc, _ = port.frame_filter(g, pr, 0, want_integral=False)
        refs.append([pr, fh, fv, fc])
    for satd in (1, 0):
        cases = bidir_cases(pkg, g, 66 + satd, 600)
        jobs = np.zeros(len(cases), pkg.BIDIR_JOB)
        for k, (j, mi, mvp0, mvp1, mv0, mv1, weight) in enumerate(cases):
            jobs[k]["bx"], jobs[k]["by"], jobs[k]["i_pixel"], jobs[k]["qp"], jobs[k]["weight"] = mi.bx, mi.by, mi.i_pixel, mi.qp, weight
            jobs[k]["flags"] = pkg.ME_MBCMP_SATD if satd else 0
            jobs[k]["mv0"], jobs[k]["mv1"], jobs[k]["mvp0"], jobs[k]["mvp1"] = mv0, mv1, mvp0, mvp1
            jobs[k]["mv_min_spel"] = [mi.mv_min_spel[0], mi.mv_min_spel[1]]
            jobs[k]["mv_max_spel"] = [mi.mv_max_spel[0], mi.mv_max_spel[1]]
        res = ctx.me_refine_bidir(fenc, frefs[0], frefs[1], jobs)
        bad, moved = [], 0
        for k, (j, mi, mvp0, mvp1, mv0, mv1, weight) in enumerate(cases):
            a0, a1, cost = port.me_refine_bidir_satd(g, pe, refs[0], refs[1], mi, mvp0, mvp1, weight, satd, mv0, mv1)
            got = (tuple(int(x) for x in res[k]["mv0"]), tuple(int(x) for x in res[k]["mv1"]), int(res[k]["cost"]))
            if got != (a0, a1, cost):
                bad.append((k, mi.i_pixel, got, (a0, a1, cost)))
            moved += (a0, a1) != (tuple(mv0), tuple(mv1))
        assert not bad, (satd, len(bad), bad[:4])
        assert moved > 300
    fenc.close()
    for f in frefs:
        f.close()
