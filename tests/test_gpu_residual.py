"""gpu: batched transforms / quantisation and the inter-macroblock residual path vs the oracle, bit-exact"""
import ctypes as C

import numpy as np
import pytest
import xo_api as X

pytestmark = pytest.mark.gpu


RIn, ROut = X.ResidIn, X.ResidOut


def _blocks(rng, n, bs):
    """checkasm-like inputs: random, zero-residual, max-difference and small-noise blocks"""
    pred = rng.integers(0, 256, (n, bs), dtype=np.uint8)
    amp = rng.choice([0, 1, 2, 5, 20, 80, 255], n)
    noise = (rng.integers(-255, 256, (n, bs)) * amp[:, None]) // 255
    fenc = np.clip(pred.astype(np.int32) + noise, 0, 255).astype(np.uint8)
    fenc[0], pred[0] = 255, 0
    fenc[1], pred[1] = 0, 255
    fenc[2] = pred[2]
    return fenc, pred


@pytest.mark.parametrize("cqm", [0, 1])
@pytest.mark.parametrize("kind", [0, 1])
def test_block_residual(pkg, ctx, port, kind, cqm):
    rng = np.random.default_rng(10 * kind + cqm)
    n, bs = 4000, 64 if kind else 16
    fenc, pred = _blocks(rng, n, bs)
    qp = rng.integers(0, 52, n).astype(np.uint8)
    qp[:52] = np.arange(52)
    cat = rng.integers(0, 2 if kind else 4, n).astype(np.uint8)
    ctx.set_quant_preset(cqm)
    got = ctx.block_residual(kind, fenc, pred, qp, cat)
    side = 8 if kind else 4
    fe_t, fd_t = np.zeros(16 * 16, np.uint8), np.zeros(32 * 16, np.uint8)
    L = port.lib
    for i in range(n):
        fe_t.reshape(16, 16)[:side, :side] = fenc[i].reshape(side, side)
        fd_t.reshape(16, 32)[:side, :side] = pred[i].reshape(side, side)
        d = np.zeros(bs, np.int16)
        (L.xo_sub8x8_dct8 if kind else L.xo_sub4x4_dct)(X._ptr(d, X.i16p), X._ptr(fe_t), X._ptr(fd_t))
        assert np.array_equal(d, got["dct"][i]), ("dct", i)
        mf, bias = np.zeros(bs, np.uint16), np.zeros(bs, np.uint16)
        dq = np.zeros(6 * bs, np.int32)
        (L.xo_quant8_tables if kind else L.xo_quant4_tables)(cqm, int(cat[i]), int(qp[i]), X._ptr(mf, X.u16p), X._ptr(bias, X.u16p))
        (L.xo_dequant8_table if kind else L.xo_dequant4_table)(cqm, int(cat[i]), X._ptr(dq, X.i32p))
        nz = (L.xo_quant_8x8 if kind else L.xo_quant_4x4)(X._ptr(d, X.i16p), X._ptr(mf, X.u16p), X._ptr(bias, X.u16p))
        assert nz == got["nz"][i] and np.array_equal(d, got["level"][i]), ("quant", i, int(qp[i]))
        if nz:
            (L.xo_dequant_8x8 if kind else L.xo_dequant_4x4)(X._ptr(d, X.i16p), X._ptr(dq, X.i32p), int(qp[i]))
            (L.xo_add8x8_idct8 if kind else L.xo_add4x4_idct)(X._ptr(fd_t), X._ptr(d, X.i16p))
        assert np.array_equal(fd_t.reshape(16, 32)[:side, :side].reshape(-1), got["recon"][i]), ("recon", i, int(qp[i]))


def test_block_dc(pkg, ctx, port):
    rng = np.random.default_rng(3)
    n = 3000
    dc = rng.integers(-4096, 4096, (n, 16)).astype(np.int16)
    dc[0], dc[1] = 4080, -4080  # max DC (checkasm.c:577-589)
    qp = rng.integers(0, 52, n).astype(np.uint8)
    cat = rng.integers(0, 4, n).astype(np.uint8)
    ctx.set_quant_preset(1)
    got = ctx.block_dc(dc, qp, cat)
    L = port.lib
    for i in range(n):
        d = dc[i].copy()
        L.xo_dct4x4dc(X._ptr(d, X.i16p))
        assert np.array_equal(d, got["fwd"][i])
        mf, bias = np.zeros(16, np.uint16), np.zeros(16, np.uint16)
        dq = np.zeros(6 * 16, np.int32)
        L.xo_quant4_tables(1, int(cat[i]), int(qp[i]), X._ptr(mf, X.u16p), X._ptr(bias, X.u16p))
        L.xo_dequant4_table(1, int(cat[i]), X._ptr(dq, X.i32p))
        nz = L.xo_quant_4x4_dc(X._ptr(d, X.i16p), int(mf[0]) >> 1, int(bias[0]) << 1)
        assert nz == got["nz"][i] and np.array_equal(d, got["level"][i])
        L.xo_idct4x4dc(X._ptr(d, X.i16p))
        L.xo_dequant_4x4_dc(X._ptr(d, X.i16p), X._ptr(dq, X.i32p), int(qp[i]))
        assert np.array_equal(d, got["deq"][i])


@pytest.mark.parametrize("cqm", [0, 1])
def test_residual_inter_frame(pkg, ctx, port, cqm):
    """whole-frame inter residual coding: every MB of a CIF frame, prediction = previous frame (zero MV)"""
    from x264_vs2008_b200 import synth
    w, h = 352, 288
    clip = synth.Clip(w, h, seed=31, noise=3)
    y1, u1, v1 = clip.yuv420(1)
    y0, u0, v0 = clip.yuv420(0)
    flags = pkg.FRAME_CHROMA
    fenc, fdec = ctx.frame(w, h, flags), ctx.frame(w, h, flags)
    ctx.set_quant_preset(cqm)
    rng = np.random.default_rng(cqm)
    mbw, mbh = w // 16, h // 16
    for rep in range(3):
        fenc.upload(y1); fenc.upload_chroma(u1, v1)
        fdec.upload(y0); fdec.upload_chroma(u0, v0)
        jobs = np.zeros(mbw * mbh, pkg.RESID_JOB)
        for i in range(len(jobs)):
            jobs[i]["mb_x"], jobs[i]["mb_y"] = i % mbw, i // mbw
            jobs[i]["qp"] = int(rng.integers(10, 45)) if rep else 26
            jobs[i]["chroma_qp"] = int(rng.integers(10, 40)) if rep else 26
            jobs[i]["flags"] = int(rng.integers(0, 4)) if rep else rep
        out = ctx.residual_inter(fenc, fdec, jobs)
        ry, ru, rv = fdec.download(pkg.PLANE_FULL)[32:32 + h, 32:32 + w], fdec.download(pkg.PLANE_CB)[16:16 + h // 2, 16:16 + w // 2], \
            fdec.download(pkg.PLANE_CR)[16:16 + h // 2, 16:16 + w // 2]
        for i, j in enumerate(jobs):
            mx, my = int(j["mb_x"]), int(j["mb_y"])
            fy = np.ascontiguousarray(y1[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16])
            fu = np.ascontiguousarray(u1[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8]); fv = np.ascontiguousarray(v1[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8])
            py = np.ascontiguousarray(y0[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16])
            pu = np.ascontiguousarray(u0[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8]); pv = np.ascontiguousarray(v0[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8])
            rin = RIn(int(j["qp"]), int(j["chroma_qp"]), int(j["flags"]) & 1, (int(j["flags"]) >> 1) & 1, cqm)
            o = ROut()
            port.lib.xo_residual_inter_mb(C.byref(rin), X._ptr(fy), X._ptr(fu), X._ptr(fv), X._ptr(py), X._ptr(pu), X._ptr(pv), C.byref(o))
            g = out[i]
            tag = (rep, i, int(j["qp"]), int(j["chroma_qp"]), int(j["flags"]))
            assert (int(g["cbp_luma"]), int(g["cbp_chroma"])) == (o.cbp_luma, o.cbp_chroma), tag
            assert list(g["nnz"]) == list(o.nnz), tag
            want_luma = np.array(o.luma8x8).reshape(-1) if rin.b_transform_8x8 else np.array(o.luma4x4)[:16].reshape(-1)
            assert np.array_equal(g["luma"], want_luma), tag
            assert np.array_equal(g["chroma_ac"], np.array(o.luma4x4)[16:24]), tag
            assert np.array_equal(g["chroma_dc"], np.array(o.chroma_dc)), tag
            assert np.array_equal(ry[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16], py), tag
            assert np.array_equal(ru[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8], pu) and np.array_equal(rv[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8], pv), tag
    fenc.close(); fdec.close()


def _intra16_modes(rng, mx, my):
    """a valid (mode16, mode_chroma) pair for the neighbours a macroblock has (analyse.c:372-440)"""
    if mx and my:
        return int(rng.integers(0, 4)), int(rng.integers(0, 4))
    if mx:
        return int(rng.choice([4, 1])), int(rng.choice([4, 1]))      # DC_LEFT, H
    if my:
        return int(rng.choice([5, 0])), int(rng.choice([5, 2]))      # DC_TOP, V (luma 0, chroma 2)
    return 6, 6


def _intra16_reference(port, jobs, cqm, y1, u1, v1, ry, ru, rv):
    """the listed macroblocks one after the other through the oracle, reading neighbours from / writing into the running reconstruction"""
    outs = []
    pads = [np.pad(a, 1) for a in (ry, ru, rv)]  # row above / column left of the frame edge: never used by a valid mode
    for j in jobs:
        mx, my = int(j["mb_x"]), int(j["mb_y"])
        f = [np.ascontiguousarray(a[my * n:my * n + n, mx * n:mx * n + n]) for a, n in ((y1, 16), (u1, 8), (v1, 8))]
        nb = []
        for a, n in zip(pads, (16, 8, 8)):
            x0, y0 = mx * n + 1, my * n + 1
            nb.append(np.concatenate([[a[y0 - 1, x0 - 1]], a[y0 - 1, x0:x0 + n], a[y0:y0 + n, x0 - 1]]).astype(np.uint8))
        rin = RIn(int(j["qp"]), int(j["chroma_qp"]), 0, (int(j["flags"]) >> 1) & 1, cqm)
        o, dc, qy, qu, qv = port.residual_intra16_mb(rin, int(j["mode16"]), int(j["mode_chroma"]), f[0], f[1], f[2], nb[0], nb[1], nb[2])
        for a, q, n in zip(pads, (qy, qu, qv), (16, 8, 8)):
            a[my * n + 1:my * n + n + 1, mx * n + 1:mx * n + n + 1] = q.reshape(n, n)
        outs.append((o, dc))
    return outs, [a[1:-1, 1:-1] for a in pads]


@pytest.mark.parametrize("size,cqm,subset", [((352, 288), 0, False), ((352, 288), 1, True), ((64, 48), 0, False), ((1920, 1080), 0, False),
                                             ((1920, 1080), 1, True)])
def test_residual_intra16_frame(pkg, ctx, port, size, cqm, subset):
    """I_16x16 macroblocks of a frame as one wavefront launch (x264_mb_encode_i16x16 + intra chroma): coefficients and reconstruction
    against the oracle run macroblock by macroblock in list order; subset: scattered intra macroblocks among final (inter) neighbours"""
    from x264_vs2008_b200 import synth
    w, h = size
    clip = synth.Clip(w, h, seed=37, noise=3)
    y1, u1, v1 = clip.yuv420(1)
    y0, u0, v0 = clip.yuv420(0)
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    hp, wp = mbh * 16, mbw * 16
    padto = lambda a, hh, ww: np.pad(a, ((0, hh - a.shape[0]), (0, ww - a.shape[1])), mode="edge")
    y1, y0 = padto(y1, hp, wp), padto(y0, hp, wp)
    u1, v1, u0, v0 = (padto(a, hp // 2, wp // 2) for a in (u1, v1, u0, v0))
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    ctx.set_quant_preset(cqm)
    rng = np.random.default_rng(17 + cqm)
    for rep in range(2):
        sel = [i for i in range(mbw * mbh) if not subset or rng.random() < 0.3]
        jobs = np.zeros(len(sel), pkg.INTRA16_JOB)
        for k, i in enumerate(sel):
            mx, my = i % mbw, i // mbw
            jobs[k]["mb_x"], jobs[k]["mb_y"] = mx, my
            jobs[k]["qp"] = int(rng.integers(12, 46)) if rep else 26
            jobs[k]["chroma_qp"] = int(rng.integers(12, 40)) if rep else 26
            jobs[k]["mode16"], jobs[k]["mode_chroma"] = _intra16_modes(rng, mx, my)
            jobs[k]["flags"] = pkg.RESID_DECIMATE if (rep and rng.random() < 0.5) else 0
        fenc.upload(y1[:h, :w]); fenc.upload_chroma(u1[:h // 2, :w // 2], v1[:h // 2, :w // 2]); fenc.expand_border_mod16()
        fdec.upload(y0[:h, :w]); fdec.upload_chroma(u0[:h // 2, :w // 2], v0[:h // 2, :w // 2]); fdec.expand_border_mod16()
        want, recon = _intra16_reference(port, jobs, cqm, y1, u1, v1, y0.copy(), u0.copy(), v0.copy())
        out = ctx.residual_intra16(fenc, fdec, jobs)
        gy = fdec.download(pkg.PLANE_FULL)[32:32 + hp, 32:32 + wp]
        gu, gv = (fdec.download(pl)[16:16 + hp // 2, 16:16 + wp // 2] for pl in (pkg.PLANE_CB, pkg.PLANE_CR))
        for k, j in enumerate(jobs):
            o, dc = want[k]
            g = out[k]
            mx, my = int(j["mb_x"]), int(j["mb_y"])
            tag = (rep, k, mx, my, int(j["qp"]), int(j["chroma_qp"]), int(j["mode16"]), int(j["mode_chroma"]), int(j["flags"]))
            assert (int(g["c"]["cbp_luma"]), int(g["c"]["cbp_chroma"])) == (o.cbp_luma, o.cbp_chroma), tag
            assert list(g["c"]["nnz"]) == list(o.nnz), tag
            assert np.array_equal(g["c"]["luma"], np.array(o.luma4x4)[:16].reshape(-1)), tag
            assert np.array_equal(g["luma_dc"], dc), tag
            assert np.array_equal(g["c"]["chroma_ac"], np.array(o.luma4x4)[16:24]), tag
            assert np.array_equal(g["c"]["chroma_dc"], np.array(o.chroma_dc)), tag
            assert np.array_equal(gy[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16], recon[0][my * 16:my * 16 + 16, mx * 16:mx * 16 + 16]), tag
        assert np.array_equal(gy, recon[0]) and np.array_equal(gu, recon[1]) and np.array_equal(gv, recon[2])
    fenc.close(); fdec.close()


@pytest.mark.parametrize("cqm", [0, 1])
def test_probe_skip_pred_in_fdec(pkg, ctx, port, cqm):
    """x264_macroblock_probe_skip, b_bidir form: tiles around the skip decision laid out as the macroblocks of a CIF frame pair"""
    import helpers
    w, h = 352, 288
    mbw, mbh = w // 16, h // 16
    cases = helpers.skip_probe_cases(70 + cqm, mbw * mbh)
    fy, py = np.zeros((h, w), np.uint8), np.zeros((h, w), np.uint8)
    fu, fv, pu, pv = (np.zeros((h // 2, w // 2), np.uint8) for _ in range(4))
    jobs = np.zeros(len(cases), pkg.SKIP_JOB)
    for i, (qp, cqp, a, b, c, d, e, f) in enumerate(cases):
        mx, my = i % mbw, i // mbw
        fy[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16], py[my * 16:my * 16 + 16, mx * 16:mx * 16 + 16] = a, d
        fu[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8], pu[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8] = b, e
        fv[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8], pv[my * 8:my * 8 + 8, mx * 8:mx * 8 + 8] = c, f
        jobs[i] = (mx, my, 0, 0, qp, cqp, pkg.SKIP_PRED_IN_FDEC, 0)
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    fenc.upload(fy); fenc.upload_chroma(fu, fv)
    fdec.upload(py); fdec.upload_chroma(pu, pv)
    ctx.set_quant_preset(cqm)
    got = ctx.probe_skip(fenc, None, fdec, jobs)
    want = np.array([port.probe_skip_mb(X.ResidIn(qp, cqp, 0, 1, cqm), a, b, c, d, e, f) for (qp, cqp, a, b, c, d, e, f) in cases], np.uint8)
    bad = np.nonzero(got != want)[0]
    assert len(bad) == 0, (bad[:8], [cases[k][:2] for k in bad[:8]])
    assert len(cases) // 5 < int(want.sum()) < 4 * len(cases) // 5
    ctx.set_quant_preset(0)
    fenc.close(); fdec.close()


def test_probe_skip_mc(pkg, ctx, port):
    """P-skip form: prediction = mc_luma / mc_chroma of the job's vector from the reference's half-pel and chroma planes, formed inside
    the kernel (and stored to fdec on request)"""
    from x264_vs2008_b200 import synth
    from helpers import padded_chroma
    w, h = 352, 288
    mbw, mbh = w // 16, h // 16
    clip = synth.Clip(w, h, seed=41, noise=1)
    (y0, u0, v0), (y1, u1, v1) = clip.yuv420(0), clip.yuv420(1)
    g = port.geometry(w, h)
    fref = ctx.frame(w, h, pkg.FRAME_HPEL | pkg.FRAME_CHROMA)
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    fref.upload(y0); fref.upload_chroma(u0, v0); fref.expand_border(); fref.filter()
    fenc.upload(y1); fenc.upload_chroma(u1, v1)
    plane = port.plane_from_picture(g, y0)
    fh, fv, fc, _ = port.frame_filter(g, plane, 0, want_integral=False)
    planes, cu, cv = [plane, fh, fv, fc], padded_chroma(g, u0), padded_chroma(g, v0)
    rng = np.random.default_rng(8)
    sent = np.full((h, w), 77, np.uint8)
    for rep in range(3):
        fdec.upload(sent); fdec.upload_chroma(sent[:h // 2, :w // 2], sent[:h // 2, :w // 2])
        jobs = np.zeros(mbw * mbh, pkg.SKIP_JOB)
        ey, eu, ev = y1.copy(), u1.copy(), v1.copy()
        preds = []
        for i in range(len(jobs)):
            mx, my = i % mbw, i // mbw
            mv = (0, 0) if rng.integers(0, 3) == 0 else (int(rng.integers(-6, 7)), int(rng.integers(-6, 7))) if rng.integers(0, 2) else \
                (int(rng.integers(-60, 61)), int(rng.integers(-60, 61)))
            jobs[i] = (mx, my, mv[0], mv[1], int(rng.integers(18, 52)), int(rng.integers(18, 52)), pkg.SKIP_STORE_PRED * int(rng.integers(0, 2)), 0)
            py = np.zeros((16, 16), np.uint8)
            arr = (X.u8p * 4)(*[X._ptr(p, X.u8p, g.origin + my * 16 * g.stride + mx * 16) for p in planes])
            port.lib.xo_mc_luma(X._ptr(py), 16, arr, g.stride, mv[0], mv[1], 16, 16)
            pc = []
            for cp in (cu, cv):
                t = np.zeros((8, 8), np.uint8)
                port.lib.xo_mc_chroma(X._ptr(t), 8, X._ptr(cp, X.u8p, (16 + my * 8) * cp.shape[1] + 16 + mx * 8), cp.shape[1], mv[0], mv[1], 8, 8)
                pc.append(t)
            preds.append((py, pc[0], pc[1]))
            if i % 4:  # most macroblocks: source = that prediction + a little noise, so the decision is close; the rest keep frame 1
                amp = int(rng.integers(0, 4))
                for dst, src, n in ((ey, py, 16), (eu, pc[0], 8), (ev, pc[1], 8)):
                    dst[my * n:my * n + n, mx * n:mx * n + n] = np.clip(src.astype(np.int32) + rng.integers(-amp, amp + 1, src.shape), 0, 255)
        fenc.upload(ey); fenc.upload_chroma(eu, ev)
        got = ctx.probe_skip(fenc, fref, fdec, jobs)
        ry, ru, rv = fdec.download(pkg.PLANE_FULL)[32:32 + h, 32:32 + w], fdec.download(pkg.PLANE_CB)[16:16 + h // 2, 16:16 + w // 2], \
            fdec.download(pkg.PLANE_CR)[16:16 + h // 2, 16:16 + w // 2]
        n1 = 0
        for i, j in enumerate(jobs):
            mx, my = int(j["mb_x"]), int(j["mb_y"])
            py, pu, pv = preds[i]
            tile = lambda a, n: np.ascontiguousarray(a[my * n:my * n + n, mx * n:mx * n + n])
            want = port.probe_skip_mb(X.ResidIn(int(j["qp"]), int(j["chroma_qp"]), 0, 1, 0), tile(ey, 16), tile(eu, 8), tile(ev, 8), py, pu, pv)
            assert int(got[i]) == want, (rep, i, int(j["mvx"]), int(j["mvy"]), int(j["qp"]), int(j["chroma_qp"]))
            n1 += want
            if int(j["flags"]) & pkg.SKIP_STORE_PRED:
                assert np.array_equal(tile(ry, 16), py) and np.array_equal(tile(ru, 8), pu) and np.array_equal(tile(rv, 8), pv), (rep, i)
            else:
                assert (tile(ry, 16) == 77).all() and (tile(ru, 8) == 77).all() and (tile(rv, 8) == 77).all(), (rep, i)
        assert len(jobs) // 8 < n1 < 7 * len(jobs) // 8, n1
    for f in (fref, fenc, fdec):
        f.close()


def test_intra16_and_upload_refusals(pkg, ctx):
    """out-of-range input is refused with -1 and a message instead of being written somewhere: a macroblock outside the frame or a mode
    number outside the reference's enums (x264_cuda_residual_intra16), a picture larger than the frame's planes (x264_cuda_frame_upload*)"""
    w, h = 64, 48
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    ctx.set_quant_preset(0)
    for bad in (dict(mb_x=4), dict(mb_y=3), dict(mode16=7), dict(mode_chroma=9), dict(mb_x=-1)):
        jobs = np.zeros(1, pkg.INTRA16_JOB)
        jobs[0]["qp"], jobs[0]["chroma_qp"], jobs[0]["mode16"], jobs[0]["mode_chroma"] = 26, 26, 6, 6
        for k, v in bad.items():
            jobs[0][k] = v
        with pytest.raises(pkg.CudaError, match="out of range"):
            ctx.residual_intra16(fenc, fdec, jobs)
    big = np.zeros((h + 17, w), np.uint8)
    with pytest.raises(pkg.CudaError, match="does not fit"):
        fenc.upload(big)
    with pytest.raises(pkg.CudaError, match="does not fit"):
        fenc.upload_chroma(np.zeros((h // 2, w // 2 + 9), np.uint8), np.zeros((h // 2, w // 2), np.uint8))
    ok = np.zeros((h, w), np.uint8)
    fenc.upload(ok)  # the frame itself still works
    fenc.close(); fdec.close()
