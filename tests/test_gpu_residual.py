"""gpu: batched transforms / quantisation and the inter-macroblock residual path vs the oracle, bit-exact"""
import ctypes as C

import numpy as np
import pytest
import xo_api as X

pytestmark = pytest.mark.gpu


class RIn(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("qp", "chroma_qp", "b_transform_8x8", "b_decimate", "cqm")]


class ROut(C.Structure):
    _fields_ = [("luma4x4", (C.c_int16 * 16) * 24), ("luma8x8", (C.c_int16 * 64) * 4), ("chroma_dc", (C.c_int16 * 4) * 2),
                ("nnz", C.c_uint8 * 27), ("pad", C.c_uint8), ("cbp_luma", C.c_int), ("cbp_chroma", C.c_int)]


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
