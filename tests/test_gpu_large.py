"""gpu: parity at the sizes BASELINE.json names (VERDICT r1 "close the untested configs"): the macroblock-batched ESA kernel on the exact
bench.py workload (every one of the 73 440 searches of a 1080p pair), and motion search (batched, small-pattern, TESA), residual coding,
skip probe, deblocking and the lookahead cost at 1080p / 3840x2160 / 7680x4320 — whole frames where the oracle finishes in seconds,
a few hundred sampled macroblocks / blocks (frame edges included) where it does not.  int16 block coordinates, 7808-byte strides and
the uint16 integral wrap are first exercised at these sizes."""
import ctypes as C

import numpy as np
import pytest
import xo_api as X
from helpers import (make_me_jobs, make_mb_jobs, mb_jobs_to_block_jobs, block_jobs_to_mis, make_deblock_info, lowres_planes, oracle_lookahead,
                     lookahead_digest)

pytestmark = pytest.mark.gpu

UHD, UHD8 = (3840, 2160), (7680, 4320)
_clips = {}


def _clip(w, h, seed=9):
    from x264_vs2008_b200 import synth
    key = (w, h, seed)
    if key not in _clips:
        _clips.clear()   # one large clip in memory at a time
        _clips[key] = synth.Clip(w, h, seed=seed)
    return _clips[key]


def _pair(pkg, ctx, port, w, h, flags_ref=0, seed=9):
    clip = _clip(w, h, seed)
    g = port.geometry(w, h)
    fenc, fref = ctx.frame(w, h, 0), ctx.frame(w, h, flags_ref)
    y1, y0 = clip.luma(1), clip.luma(0)
    fenc.upload(y1); fenc.expand_border()
    fref.upload(y0); fref.expand_border()
    if flags_ref:
        fref.filter()
    return g, fenc, fref, port.plane_from_picture(g, y1), port.plane_from_picture(g, y0)


def test_bench_workload_every_search_1080p(pkg, ctx, port):
    """BASELINE config 2 as bench.py runs it: all 8160 macroblocks x 9 partition searches of the 1080p pair, x264_cuda_me_search_mb vs
    the reference's own x264_me_search_ref where oracle/_ref exists (else the port) — all 73 440 (mv, cost) results"""
    import bench
    w, h = bench.W, bench.H
    g, fenc, fref, pe, pr = _pair(pkg, ctx, port, w, h, seed=1)
    jobs = bench.build_jobs(pkg, g.mb_width, g.mb_height)
    mbjobs = bench.to_mb_jobs(pkg, jobs, g.mb_width, g.mb_height)
    ctx.set_cost_mv(bench.QP)
    res = ctx.me_search_mb(fenc, fref, bench.ME_RANGE, mbjobs)
    got = np.stack([res["part"]["bmx"].reshape(-1), res["part"]["bmy"].reshape(-1), res["part"]["bcost"].reshape(-1)], 1).astype(np.int64)
    mis = block_jobs_to_mis(jobs, bench.ME_RANGE)
    outs = port.me_search_fpel_batch(g, pe, pr, None, mis)
    want = np.array([(r.bmx, r.bmy, r.bcost) for r in outs], np.int64)
    bad = np.nonzero((got != want).any(1))[0]
    assert len(bad) == 0, (len(bad), bad[:5], got[bad[:5]], want[bad[:5]])
    seeds = np.stack([res["part"]["seed_mx"].reshape(-1), res["part"]["seed_my"].reshape(-1), res["part"]["seed_cost"].reshape(-1)], 1).astype(np.int64)
    assert np.array_equal(seeds, np.array([(r.seed_mx, r.seed_my, r.seed_cost) for r in outs], np.int64))
    if X.have_ref():
        # the unmodified reference (x264_me_search_ref with its ADS-accelerated loop, subme 1) only exposes m->mv / m->cost: derive the same
        # from the device's full-pel result (me.c:603-620: cost gets the vector cost back when the winner is the rounded prediction)
        ref = X.ref()
        _, _, _, integ = ref.frame_filter(g, pr, 0)
        routs = ref.me_search_fpel_batch(g, pe, pr, integ, mis)
        tab = pkg.host_cost_mv(bench.QP).astype(np.int64)
        c0 = 2 * 4 * 2048
        mvp = jobs["mvp"].astype(np.int64)
        lo, hi = jobs["mv_min_fpel"].astype(np.int64), jobs["mv_max_fpel"].astype(np.int64)
        pm = (np.clip(mvp, lo * 4, hi * 4) + 2) >> 2
        cost_mv = tab[c0 + 4 * got[:, 0] - mvp[:, 0]] + tab[c0 + 4 * got[:, 1] - mvp[:, 1]]
        cost = got[:, 2] + np.where((got[:, 0] == pm[:, 0]) & (got[:, 1] == pm[:, 1]), cost_mv, 0)
        mine = np.stack([4 * got[:, 0], 4 * got[:, 1], cost, cost_mv], 1)
        theirs = np.array([(r.mv[0], r.mv[1], r.cost, r.cost_mv) for r in routs], np.int64)
        bad = np.nonzero((mine != theirs).any(1))[0]
        assert len(bad) == 0, ("vs reference", len(bad), bad[:5], mine[bad[:5]], theirs[bad[:5]])
    assert len(got) == 73440
    fenc.close(); fref.close()


@pytest.mark.parametrize("size", [UHD, UHD8])
def test_mb_search_large(pkg, ctx, port, size):
    w, h = size
    g, fenc, fref, pe, pr = _pair(pkg, ctx, port, w, h)
    mbjobs = make_mb_jobs(pkg, g, seed=w, n=320, qp=(12, 26, 38))
    # make sure the far corner and both far edges are among the sampled macroblocks
    for k, (mx, my) in enumerate([(g.mb_width - 1, g.mb_height - 1), (g.mb_width - 1, 0), (0, g.mb_height - 1), (g.mb_width - 2, g.mb_height // 2)]):
        mn, mxl, _, _ = X.mv_limits_fpel(g, mx, my)
        mbjobs[k]["mb_x"], mbjobs[k]["mb_y"], mbjobs[k]["mv_min_fpel"], mbjobs[k]["mv_max_fpel"] = mx, my, mn, mxl
    res = ctx.me_search_mb(fenc, fref, 16, mbjobs)
    bjobs, idx = mb_jobs_to_block_jobs(pkg, mbjobs)
    outs = port.me_search_fpel_batch(g, pe, pr, None, block_jobs_to_mis(bjobs, 16))
    bres = ctx.me_search(fenc, fref, 16, bjobs)
    bad = []
    for k, (i, p) in enumerate(idx):
        r, o, b = res[i]["part"][p], outs[k], bres[k]
        got = (int(r["bmx"]), int(r["bmy"]), int(r["bcost"]), int(r["seed_mx"]), int(r["seed_my"]), int(r["seed_cost"]))
        if got != (o.bmx, o.bmy, o.bcost, o.seed_mx, o.seed_my, o.seed_cost) or got[:3] != (int(b["bmx"]), int(b["bmy"]), int(b["bcost"])):
            bad.append((i, p, got, (o.bmx, o.bmy, o.bcost)))
    assert not bad, (len(bad), bad[:5])
    fenc.close(); fref.close()


@pytest.mark.parametrize("size", [UHD, UHD8])
def test_small_and_tesa_search_large(pkg, ctx, port, size):
    """HEX + subme 7 and the complete --me tesa search (ADS on the wrapped uint16 integral image) on sampled blocks of a 4K / 8K pair"""
    w, h = size
    g, fenc, fref, pe, pr = _pair(pkg, ctx, port, w, h, flags_ref=pkg.FRAME_HPEL | pkg.FRAME_INTEGRAL | pkg.FRAME_INTEGRAL4)
    fh, fv, fc, integ = port.frame_filter(g, pr, 1)
    planes = [pr, fh, fv, fc]
    for method, tesa, subme, n in ((X.ME_HEX, False, 7, 240), (X.ME_TESA, True, 2, 240)):
        jobs, mis = make_me_jobs(pkg, g, seed=w + subme, n=n, me_range=16, qp=(12, 26, 38), pixels=(0, 1, 2, 3, 4, 5, 6), mvp_spread=40, tesa=tesa, fpel_satd=tesa)
        for j, mi in zip(jobs, mis):
            j["mv_min_spel"] = [mi.mv_min_spel[0], mi.mv_min_spel[1]]
            j["mv_max_spel"] = [mi.mv_max_spel[0], mi.mv_max_spel[1]]
        jobs["flags"] = pkg.ME_MBCMP_SATD | (pkg.ME_FPEL_SATD if tesa else 0)
        res = ctx.me_search_small(fenc, fref, pkg.ME_METHOD_TESA if tesa else method, 16, subme, jobs)
        bad = []
        for i, mi in enumerate(mis):
            mi.me_method = method
            mi.b_sub8x8 = 1
            o = port.me_search_subpel(g, pe, planes, integ, mi, subme, 1)
            got = (int(res[i]["mv"][0]), int(res[i]["mv"][1]), int(res[i]["cost"]), int(res[i]["cost_mv"]), int(res[i]["bmx"]), int(res[i]["bmy"]))
            if got != (o.mv[0], o.mv[1], o.cost, o.cost_mv, o.bmx, o.bmy):
                bad.append((i, mi.i_pixel, got, (o.mv[0], o.mv[1], o.cost, o.cost_mv, o.bmx, o.bmy)))
        assert not bad, (method, len(bad), bad[:4])
    fenc.close(); fref.close()


class RIn(C.Structure):
    _fields_ = [("qp", C.c_int), ("chroma_qp", C.c_int), ("b_transform_8x8", C.c_int), ("b_decimate", C.c_int), ("cqm", C.c_int)]


class ROut(C.Structure):
    _fields_ = [("luma4x4", (C.c_int16 * 16) * 24), ("luma8x8", (C.c_int16 * 64) * 4), ("chroma_dc", (C.c_int16 * 4) * 2), ("nnz", C.c_uint8 * 27),
                ("pad", C.c_uint8), ("cbp_luma", C.c_int), ("cbp_chroma", C.c_int)]


@pytest.mark.parametrize("size,n_sample", [((1920, 1080), 0), (UHD, 500), (UHD8, 500)])
def test_residual_inter_large(pkg, ctx, port, size, n_sample):
    """inter residual coding: EVERY macroblock of a 1080p frame (config 3's size), 500 sampled ones at 4K / 8K; prediction = previous frame"""
    w, h = size
    clip = _clip(w, h, 31)
    y1, u1, v1 = clip.yuv420(1)
    y0, u0, v0 = clip.yuv420(0)
    g = port.geometry(w, h)
    mbw, mbh = g.mb_width, h // 16   # whole macroblocks of the picture
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    ctx.set_quant_preset(0)
    rng = np.random.default_rng(w)
    ids = np.arange(mbw * mbh) if not n_sample else np.unique(np.concatenate([rng.choice(mbw * mbh, n_sample, replace=False), [0, mbw - 1, mbw * mbh - 1, mbw * (mbh - 1)]]))
    fenc.upload(y1); fenc.upload_chroma(u1, v1)
    fdec.upload(y0); fdec.upload_chroma(u0, v0)
    jobs = np.zeros(len(ids), pkg.RESID_JOB)
    jobs["mb_x"], jobs["mb_y"] = ids % mbw, ids // mbw
    jobs["qp"], jobs["chroma_qp"], jobs["flags"] = rng.integers(10, 45, len(ids)), rng.integers(10, 40, len(ids)), rng.integers(0, 4, len(ids))
    out = ctx.residual_inter(fenc, fdec, jobs)
    ry, ru, rv = fdec.download(pkg.PLANE_FULL)[32:32 + h, 32:32 + w], fdec.download(pkg.PLANE_CB)[16:16 + h // 2, 16:16 + w // 2], \
        fdec.download(pkg.PLANE_CR)[16:16 + h // 2, 16:16 + w // 2]
    for i, j in enumerate(jobs):
        mx, my = int(j["mb_x"]), int(j["mb_y"])
        t = lambda a, n: np.ascontiguousarray(a[my * n:my * n + n, mx * n:mx * n + n])
        fy, fu, fv, py, pu, pv = t(y1, 16), t(u1, 8), t(v1, 8), t(y0, 16), t(u0, 8), t(v0, 8)
        rin, o = RIn(int(j["qp"]), int(j["chroma_qp"]), int(j["flags"]) & 1, (int(j["flags"]) >> 1) & 1, 0), ROut()
        port.lib.xo_residual_inter_mb(C.byref(rin), X._ptr(fy), X._ptr(fu), X._ptr(fv), X._ptr(py), X._ptr(pu), X._ptr(pv), C.byref(o))
        gg = out[i]
        tag = (i, mx, my, int(j["qp"]), int(j["chroma_qp"]), int(j["flags"]))
        assert (int(gg["cbp_luma"]), int(gg["cbp_chroma"])) == (o.cbp_luma, o.cbp_chroma) and list(gg["nnz"]) == list(o.nnz), tag
        want_luma = np.array(o.luma8x8).reshape(-1) if rin.b_transform_8x8 else np.array(o.luma4x4)[:16].reshape(-1)
        assert np.array_equal(gg["luma"], want_luma) and np.array_equal(gg["chroma_ac"], np.array(o.luma4x4)[16:24]), tag
        assert np.array_equal(gg["chroma_dc"], np.array(o.chroma_dc)), tag
        assert np.array_equal(t(ry, 16), py) and np.array_equal(t(ru, 8), pu) and np.array_equal(t(rv, 8), pv), tag
    fenc.close(); fdec.close()


@pytest.mark.parametrize("size,n_sample", [((1920, 1080), 0), (UHD, 600), (UHD8, 600)])
def test_probe_skip_large(pkg, ctx, port, size, n_sample):
    """x264_macroblock_probe_skip (prediction in fdec): every macroblock of 1080p, sampled ones at 4K / 8K; source = prediction + noise
    of a few amplitudes so that the decisions are close"""
    w, h = size
    clip = _clip(w, h, 41)
    y0, u0, v0 = clip.yuv420(0)
    g = port.geometry(w, h)
    mbw, mbh = g.mb_width, h // 16
    rng = np.random.default_rng(h)
    ids = np.arange(mbw * mbh) if not n_sample else np.unique(np.concatenate([rng.choice(mbw * mbh, n_sample, replace=False), [0, mbw - 1, mbw * mbh - 1]]))
    ey, eu, ev = y0.copy(), u0.copy(), v0.copy()
    jobs = np.zeros(len(ids), pkg.SKIP_JOB)
    for k, i in enumerate(ids):
        mx, my = int(i % mbw), int(i // mbw)
        amp = int(rng.integers(0, 5))
        for dst, n in ((ey, 16), (eu, 8), (ev, 8)):
            blk = dst[my * n:my * n + n, mx * n:mx * n + n]
            blk[...] = np.clip(blk.astype(np.int32) + rng.integers(-amp, amp + 1, blk.shape), 0, 255)
        jobs[k] = (mx, my, 0, 0, int(rng.integers(14, 48)), int(rng.integers(14, 48)), pkg.SKIP_PRED_IN_FDEC, 0)
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    fenc.upload(ey); fenc.upload_chroma(eu, ev)
    fdec.upload(y0); fdec.upload_chroma(u0, v0)
    ctx.set_quant_preset(0)
    got = ctx.probe_skip(fenc, None, fdec, jobs)
    want = []
    for j in jobs:
        mx, my = int(j["mb_x"]), int(j["mb_y"])
        t = lambda a, n: np.ascontiguousarray(a[my * n:my * n + n, mx * n:mx * n + n])
        want.append(port.probe_skip_mb(X.ResidIn(int(j["qp"]), int(j["chroma_qp"]), 0, 1, 0), t(ey, 16), t(eu, 8), t(ev, 8), t(y0, 16), t(u0, 8), t(v0, 8)))
    want = np.array(want, np.uint8)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (len(bad), bad[:8])
    assert len(jobs) // 16 < int(want.sum()) < 15 * len(jobs) // 16, int(want.sum())  # both outcomes well represented (content-dependent: a wide band)
    fenc.close(); fdec.close()


@pytest.mark.parametrize("size,kw", [(UHD, dict()), (UHD, dict(chaos=True, slice_b=1)), (UHD8, dict()), (UHD8, dict(chaos=True, slice_b=1, cavlc_8x8dct=1, alpha=4, beta=-2))])
def test_frame_deblock_large(pkg, ctx, port, size, kw):
    from test_gpu_deblock import run_both
    w, h = size
    g = port.geometry(w, h)
    run_both(pkg, ctx, port, w, h, make_deblock_info(g, seed=300 + w, **kw), 3)


@pytest.mark.parametrize("size", [UHD])
def test_lowres_frame_cost_large(pkg, ctx, port, size):
    """the lookahead's whole evaluation schedule (I, P, P dist 2, B, cached) on a 3840x2160 window: 32 400 blocks per evaluation"""
    from test_gpu_lookahead import device_lookahead
    w, h = size
    g = port.geometry(w, h)
    clip = _clip(w, h, 31)
    planes = lowres_planes(port, g, clip, 3)
    want = oracle_lookahead(port, g, planes, X.ME_HEX, 16, 1, 1)
    got = device_lookahead(pkg, ctx, g, clip, X.ME_HEX, 16, 1, 1)
    sa, xa = lookahead_digest(want, g)
    sb, xb = lookahead_digest(got, g)
    for i in range(len(sa)):
        bad = np.nonzero(xa[i] != xb[i])[0]
        assert len(bad) == 0, (i, len(bad), bad[:8])
        assert np.array_equal(sa[i], sb[i]), (i, sa[i], sb[i])


def test_lowres_p_cost_8k(pkg, ctx, port):
    """one P cost (and the intra costs it needs) of a 7680x4320 pair: 129 600 blocks in one wavefront launch"""
    w, h = UHD8
    g = port.geometry(w, h)
    clip = _clip(w, h, 31)
    planes = lowres_planes(port, g, clip, 2)
    n = g.mb_width * g.mb_height
    st = {"mvs0": np.zeros((n, 2), np.int16), "costs0": np.zeros(n, np.int32), "mvs1": np.zeros((n, 2), np.int16), "costs1": np.zeros(n, np.int32),
          "intra": np.zeros(n, np.uint16), "ref1_mvs": np.zeros((n, 2), np.int16)}
    o = port.lowres_frame_cost(g, planes[1], planes[0], planes[1], 0, 1, 1, st, do_search=(1, 0), b_intra_calculated=0)
    frames = []
    for i in range(2):
        f = ctx.frame(w, h, pkg.FRAME_LOWRES)
        f.upload(clip.luma(i)); f.expand_border(); f.init_lowres(); f.lookahead_alloc(1)
        frames.append(f)
    score, imbs, isum = ctx.lowres_frame_cost(frames[1], frames[0], frames[1], 0, 1, 1, me_method=X.ME_HEX, me_range=16, flags=pkg.ME_MBCMP_SATD,
                                              do_search=(1, 0), b_intra_calculated=0)
    assert (score, imbs, isum) == (o.score, o.intra_mbs, o.intra_cost_sum)
    mv, cost, _ = frames[1].lookahead_get(0, 0)
    m = np.zeros((g.mb_height, g.mb_width), bool)
    m[1:-1, 1:-1] = True
    m = m.ravel()
    assert np.array_equal(mv[m], st["mvs0"][m]) and np.array_equal(cost[m], st["costs0"][m])
    for f in frames:
        f.close()
