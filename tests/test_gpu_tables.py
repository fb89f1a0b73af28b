"""gpu: the reference's plugin tables filled by x264_*_init_cuda — checkasm-style (S/tools/checkasm.c): every overridden
entry is called through its C function pointer on host buffers and compared with the oracle."""
import ctypes as C

import numpy as np
import pytest
import xo_api as X

pytestmark = pytest.mark.gpu

u8p, i16p, u16p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_int16), C.POINTER(C.c_uint16), C.POINTER(C.c_int)
CMP = C.CFUNCTYPE(C.c_int, u8p, C.c_int, u8p, C.c_int)
CMP3 = C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, C.c_int, i32p)
CMP4 = C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, u8p, C.c_int, i32p)
VP = C.c_void_p
INTRA3 = C.CFUNCTYPE(None, u8p, u8p, i32p)


class PixelTable(C.Structure):
    _fields_ = [("sad", CMP * 7), ("ssd", CMP * 7), ("satd", CMP * 7), ("ssim", CMP * 7), ("sa8d", CMP * 4), ("mbcmp", CMP * 7),
                ("mbcmp_unaligned", CMP * 7), ("fpelcmp", CMP * 7), ("fpelcmp_x3", CMP3 * 7), ("fpelcmp_x4", CMP4 * 7),
                ("sad_aligned", CMP * 7), ("var", C.CFUNCTYPE(C.c_int, u8p, C.c_int) * 4), ("hadamard_ac", C.CFUNCTYPE(C.c_uint64, u8p, C.c_int) * 4),
                ("ssim_4x4x2_core", C.CFUNCTYPE(None, u8p, C.c_int, u8p, C.c_int, i32p)), ("ssim_end4", VP),
                ("sad_x3", CMP3 * 7), ("sad_x4", CMP4 * 7), ("satd_x3", CMP3 * 7), ("satd_x4", CMP4 * 7),
                ("ads", C.CFUNCTYPE(C.c_int, i32p, u16p, C.c_int, u16p, i16p, C.c_int, C.c_int) * 7),
                ("intra_mbcmp_x3_16x16", INTRA3), ("intra_satd_x3_16x16", INTRA3), ("intra_sad_x3_16x16", INTRA3), ("intra_satd_x3_8x8c", INTRA3),
                ("intra_satd_x3_4x4", VP), ("intra_sa8d_x3_8x8", VP)]


class DctTable(C.Structure):
    _fields_ = [(n, t) for n, t in [
        ("sub4x4_dct", C.CFUNCTYPE(None, i16p, u8p, u8p)), ("add4x4_idct", C.CFUNCTYPE(None, u8p, i16p)),
        ("sub8x8_dct", C.CFUNCTYPE(None, i16p, u8p, u8p)), ("add8x8_idct", C.CFUNCTYPE(None, u8p, i16p)),
        ("add8x8_idct_dc", C.CFUNCTYPE(None, u8p, i16p)), ("sub16x16_dct", C.CFUNCTYPE(None, i16p, u8p, u8p)),
        ("add16x16_idct", C.CFUNCTYPE(None, u8p, i16p)), ("add16x16_idct_dc", C.CFUNCTYPE(None, u8p, i16p)),
        ("sub8x8_dct8", C.CFUNCTYPE(None, i16p, u8p, u8p)), ("add8x8_idct8", C.CFUNCTYPE(None, u8p, i16p)),
        ("sub16x16_dct8", C.CFUNCTYPE(None, i16p, u8p, u8p)), ("add16x16_idct8", C.CFUNCTYPE(None, u8p, i16p)),
        ("dct4x4dc", C.CFUNCTYPE(None, i16p)), ("idct4x4dc", C.CFUNCTYPE(None, i16p))]]


class QuantTable(C.Structure):
    _fields_ = [("quant_8x8", C.CFUNCTYPE(C.c_int, i16p, u16p, u16p)), ("quant_4x4", C.CFUNCTYPE(C.c_int, i16p, u16p, u16p)),
                ("quant_4x4_dc", C.CFUNCTYPE(C.c_int, i16p, C.c_int, C.c_int)), ("quant_2x2_dc", C.CFUNCTYPE(C.c_int, i16p, C.c_int, C.c_int)),
                ("dequant_8x8", C.CFUNCTYPE(None, i16p, i32p, C.c_int)), ("dequant_4x4", C.CFUNCTYPE(None, i16p, i32p, C.c_int)),
                ("dequant_4x4_dc", C.CFUNCTYPE(None, i16p, i32p, C.c_int)), ("rest", VP * 15)]


class McTable(C.Structure):
    _fields_ = [("mc_luma", C.CFUNCTYPE(None, u8p, C.c_int, C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)),
                ("get_ref", C.CFUNCTYPE(u8p, u8p, C.POINTER(C.c_int), C.POINTER(u8p), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)),
                ("mc_chroma", C.CFUNCTYPE(None, u8p, C.c_int, u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int)),
                ("avg", C.CFUNCTYPE(None, u8p, C.c_int, u8p, C.c_int, u8p, C.c_int, C.c_int) * 10), ("copy", VP * 7), ("copy_16x16_unaligned", VP), ("plane_copy", VP),
                ("hpel_filter", C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, i16p)),
                ("prefetch_fenc", VP), ("prefetch_ref", VP), ("memcpy_aligned", VP), ("memzero_aligned", VP),
                ("integral", VP * 4),
                ("frame_init_lowres_core", C.CFUNCTYPE(None, u8p, u8p, u8p, u8p, u8p, C.c_int, C.c_int, C.c_int, C.c_int))]


def P(a, off=0, ty=u8p):
    return C.cast(a.ctypes.data + off * a.itemsize, ty)


@pytest.fixture(scope="module")
def L(pkg, ctx):
    return pkg.lib()


def test_pixel_table(pkg, L, port):
    t = PixelTable()
    assert L.x264_pixel_init_cuda(C.byref(t)) == 0
    assert not t.ssim_end4 and not t.intra_satd_x3_4x4 and not t.intra_sa8d_x3_8x8 and not t.var[1] and not t.var[2]  # left to the C table (or NULL there too)
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    b = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    fenc = np.ascontiguousarray(rng.integers(0, 256, (16, 16), dtype=np.uint8))
    for ip in range(7):
        for name, metric in (("sad", X.SAD), ("sad_aligned", X.SAD), ("ssd", X.SSD), ("satd", X.SATD)):
            for _ in range(6):
                oa, ob = int(rng.integers(0, 32)) * 64 + int(rng.integers(0, 32)), int(rng.integers(0, 32)) * 64 + int(rng.integers(0, 32))
                assert getattr(t, name)[ip](P(a, oa), 64, P(b, ob), 64) == port.pixel_cmp(metric, ip, a, 64, b, 64, oa, ob), (name, ip)
        offs = [int(rng.integers(0, 32)) * 64 + int(rng.integers(0, 32)) for _ in range(4)]
        for name, metric in (("sad", X.SAD), ("satd", X.SATD)):
            s3, s4 = (C.c_int * 3)(), (C.c_int * 4)()
            getattr(t, name + "_x3")[ip](P(fenc), P(b, offs[0]), P(b, offs[1]), P(b, offs[2]), 64, s3)
            getattr(t, name + "_x4")[ip](P(fenc), P(b, offs[0]), P(b, offs[1]), P(b, offs[2]), P(b, offs[3]), 64, s4)
            want = [port.pixel_cmp(metric, ip, fenc, 16, b, 64, 0, o) for o in offs]
            assert list(s3) == want[:3] and list(s4) == want, (name, ip)
    for ip in (0, 3):
        assert t.sa8d[ip](P(a, 65), 64, P(b, 130), 64) == port.pixel_cmp(X.SA8D, ip, a, 64, b, 64, 65, 130)
    assert not t.sa8d[1] and not t.sa8d[2]
    # ads: checkasm.c:433-464
    for ip in range(7):
        for _ in range(25):
            sums = rng.integers(0, 1 << 14, 72, dtype=np.uint16)
            dc = (C.c_int * 4)(*[int(x) for x in rng.integers(0, 1 << 14, 4)])
            cost = rng.integers(0, 1 << 10, 32, dtype=np.uint16)
            thresh = int(rng.integers(0, 1 << 15))
            m1, m2 = np.zeros(32, np.int16), np.zeros(32, np.int16)
            n1 = t.ads[ip](dc, P(sums, 0, u16p), 32, P(cost, 0, u16p), P(m1, 0, i16p), 28, thresh)
            n2 = port.lib.xo_pixel_ads(ip, dc, X._ptr(sums, X.u16p), 32, X._ptr(cost, X.u16p), X._ptr(m2, X.i16p), 28, thresh)
            assert n1 == n2 and np.array_equal(m1[:n1], m2[:n2]), ip


def test_stat_entries(pkg, L, port):
    """var, hadamard_ac, ssim_4x4x2_core (checkasm.c:330-378, :409-431) incl. flat and extreme blocks"""
    t = PixelTable()
    assert L.x264_pixel_init_cuda(C.byref(t)) == 0
    rng = np.random.default_rng(12)
    for trial in range(40):
        a = rng.integers(0, 256, (32, 64), dtype=np.uint8) if trial % 4 else np.full((32, 64), 255 * (trial % 8 == 0), np.uint8)
        b = rng.integers(0, 256, (32, 64), dtype=np.uint8)
        if trial % 5 == 1:
            a[:] = ((np.add.outer(np.arange(32), np.arange(64)) & 1) * 255).astype(np.uint8)
        off = int(rng.integers(0, 16)) * 64 + int(rng.integers(0, 40))
        for ip in (0, 3):
            assert t.var[ip](P(a, off), 64) == port.lib.xo_pixel_var(ip, X._ptr(a, X.u8p, off), 64), ("var", ip, trial)
        for ip in range(4):
            assert t.hadamard_ac[ip](P(a, off), 64) == port.lib.xo_pixel_hadamard_ac(ip, X._ptr(a, X.u8p, off), 64), ("hadamard_ac", ip, trial)
        sums = (C.c_int * 8)()
        t.ssim_4x4x2_core(P(a, off), 64, P(b, off), 64, sums)
        want = np.zeros((1, 2, 4), np.int32)
        port.lib.xo_frame_ssim_sums(X._ptr(a, X.u8p, off), 64, X._ptr(b, X.u8p, off), 64, 8, 4, X._ptr(want, X.i32p))
        assert list(sums) == list(want.reshape(-1)), ("ssim", trial)


def test_intra_x3_entries(pkg, L, port):
    """intra_{satd,sad,mbcmp}_x3_16x16 / intra_satd_x3_8x8c the way checkasm tests them (S/tools/checkasm.c:380-407): against
    predict + mbcmp of the first three modes, on an FDEC_STRIDE tile with the neighbours in place"""
    import helpers
    t = PixelTable()
    assert L.x264_pixel_init_cuda(C.byref(t)) == 0
    for i, (nbr, lam, satd, sb, fy, fu, fv, nby, nbu, nbv) in enumerate(helpers.intra_cases(95, 60)):
        for n, fenc, nb, entries in ((16, fy, nby, (("intra_satd_x3_16x16", X.SATD), ("intra_sad_x3_16x16", X.SAD), ("intra_mbcmp_x3_16x16", X.SATD))),
                                     (8, fu, nbu, (("intra_satd_x3_8x8c", X.SATD),))):
            fe = np.zeros((16, 16), np.uint8); fe[:n, :n] = fenc
            tile = np.full((17, 32), 0x55, np.uint8)   # block origin at (1, 16): row 0 is the row above, column 15 the left column
            tile[0, 15], tile[0, 16:16 + n], tile[1:1 + n, 15] = nb[0], nb[1:1 + n], nb[1 + n:1 + 2 * n]
            for name, metric in entries:
                res = (C.c_int * 3)()
                getattr(t, name)(P(fe), P(tile, 32 + 16), res)
                want = []
                for m in range(3):
                    pred = np.zeros((16, 16), np.uint8); pred[:n, :n] = port.predict(n == 8, m, nb)
                    want.append(port.pixel_cmp(metric, 0 if n == 16 else 3, pred, 16, fe, 16, 0, 0))
                assert list(res) == want, (i, name)


def test_dct_and_quant_tables(pkg, L, port):
    d, q = DctTable(), QuantTable()
    assert L.x264_dct_init_cuda(C.byref(d)) == 0 and L.x264_quant_init_cuda(C.byref(q)) == 0
    rng = np.random.default_rng(2)
    o = port.lib
    for trial in range(12):
        fe = rng.integers(0, 256, 16 * 16, dtype=np.uint8)
        fd = rng.integers(0, 256, 32 * 16, dtype=np.uint8)
        # forward transforms: 4x4-family and 8x8-family over 16x16
        got = np.zeros(256, np.int16)
        d.sub16x16_dct(P(got, 0, i16p), P(fe), P(fd))
        want = np.zeros(256, np.int16)
        for b in range(16):
            x = ((b >> 2) & 1) * 8 + (b & 1) * 4
            y = (b >> 3) * 8 + ((b >> 1) & 1) * 4
            o.xo_sub4x4_dct(X._ptr(want, X.i16p, b * 16), X._ptr(fe, X.u8p, y * 16 + x), X._ptr(fd, X.u8p, y * 32 + x))
        assert np.array_equal(got, want)
        g1 = np.zeros(16, np.int16); d.sub4x4_dct(P(g1, 0, i16p), P(fe), P(fd)); assert np.array_equal(g1, want[:16])
        g4 = np.zeros(64, np.int16); d.sub8x8_dct(P(g4, 0, i16p), P(fe), P(fd)); assert np.array_equal(g4, want[:64])
        got8, want8 = np.zeros(256, np.int16), np.zeros(256, np.int16)
        d.sub16x16_dct8(P(got8, 0, i16p), P(fe), P(fd))
        for b in range(4):
            o.xo_sub8x8_dct8(X._ptr(want8, X.i16p, b * 64), X._ptr(fe, X.u8p, (b >> 1) * 8 * 16 + (b & 1) * 8), X._ptr(fd, X.u8p, (b >> 1) * 8 * 32 + (b & 1) * 8))
        assert np.array_equal(got8, want8)
        g8 = np.zeros(64, np.int16); d.sub8x8_dct8(P(g8, 0, i16p), P(fe), P(fd)); assert np.array_equal(g8, want8[:64])
        # quant / dequant with real tables, then inverse transforms
        qp, cqm = int(rng.integers(0, 52)), int(rng.integers(0, 2))
        mf, bias, dq = np.zeros(16, np.uint16), np.zeros(16, np.uint16), np.zeros(96, np.int32)
        o.xo_quant4_tables(cqm, 1, qp, X._ptr(mf, X.u16p), X._ptr(bias, X.u16p)); o.xo_dequant4_table(cqm, 1, X._ptr(dq, X.i32p))
        a, b2 = want[:16].copy(), want[:16].copy()
        assert q.quant_4x4(P(a, 0, i16p), P(mf, 0, u16p), P(bias, 0, u16p)) == o.xo_quant_4x4(X._ptr(b2, X.i16p), X._ptr(mf, X.u16p), X._ptr(bias, X.u16p))
        assert np.array_equal(a, b2)
        q.dequant_4x4(P(a, 0, i16p), P(dq, 0, i32p), qp); o.xo_dequant_4x4(X._ptr(b2, X.i16p), X._ptr(dq, X.i32p), qp)
        assert np.array_equal(a, b2)
        r1, r2 = fd.copy(), fd.copy()
        d.add4x4_idct(P(r1), P(a, 0, i16p)); o.xo_add4x4_idct(X._ptr(r2), X._ptr(b2, X.i16p))
        assert np.array_equal(r1, r2)
        mf8, bias8, dq8 = np.zeros(64, np.uint16), np.zeros(64, np.uint16), np.zeros(384, np.int32)
        o.xo_quant8_tables(cqm, 1, qp, X._ptr(mf8, X.u16p), X._ptr(bias8, X.u16p)); o.xo_dequant8_table(cqm, 1, X._ptr(dq8, X.i32p))
        a, b2 = want8[:64].copy(), want8[:64].copy()
        assert q.quant_8x8(P(a, 0, i16p), P(mf8, 0, u16p), P(bias8, 0, u16p)) == o.xo_quant_8x8(X._ptr(b2, X.i16p), X._ptr(mf8, X.u16p), X._ptr(bias8, X.u16p))
        assert np.array_equal(a, b2)
        q.dequant_8x8(P(a, 0, i16p), P(dq8, 0, i32p), qp); o.xo_dequant_8x8(X._ptr(b2, X.i16p), X._ptr(dq8, X.i32p), qp)
        assert np.array_equal(a, b2)
        r1, r2 = fd.copy(), fd.copy()
        d.add8x8_idct8(P(r1), P(a.copy(), 0, i16p)); o.xo_add8x8_idct8(X._ptr(r2), X._ptr(b2.copy(), X.i16p))
        assert np.array_equal(r1, r2)
        # 16x16 inverse (4x4 family) on the quantised+dequantised full macroblock
        full = want.copy()
        for b in range(16):
            o.xo_quant_4x4(X._ptr(full, X.i16p, b * 16), X._ptr(mf, X.u16p), X._ptr(bias, X.u16p))
            o.xo_dequant_4x4(X._ptr(full, X.i16p, b * 16), X._ptr(dq, X.i32p), qp)
        r1, r2 = fd.copy(), fd.copy()
        d.add16x16_idct(P(r1), P(full, 0, i16p))
        for b in range(16):
            x = ((b >> 2) & 1) * 8 + (b & 1) * 4
            y = (b >> 3) * 8 + ((b >> 1) & 1) * 4
            o.xo_add4x4_idct(X._ptr(r2, X.u8p, y * 32 + x), X._ptr(full.copy(), X.i16p, b * 16))
        assert np.array_equal(r1, r2)
        # DC paths
        dc = rng.integers(-4096, 4096, 16).astype(np.int16)
        a, b2 = dc.copy(), dc.copy()
        d.dct4x4dc(P(a, 0, i16p)); o.xo_dct4x4dc(X._ptr(b2, X.i16p)); assert np.array_equal(a, b2)
        m, bs = int(mf[0]) >> 1, int(bias[0]) << 1
        assert q.quant_4x4_dc(P(a, 0, i16p), m, bs) == o.xo_quant_4x4_dc(X._ptr(b2, X.i16p), m, bs) and np.array_equal(a, b2)
        a2, b3 = a[:4].copy(), a[:4].copy()
        assert q.quant_2x2_dc(P(a2, 0, i16p), m, bs) == o.xo_quant_2x2_dc(X._ptr(b3, X.i16p), m, bs) and np.array_equal(a2, b3)
        d.idct4x4dc(P(a, 0, i16p)); o.xo_idct4x4dc(X._ptr(b2, X.i16p)); assert np.array_equal(a, b2)
        q.dequant_4x4_dc(P(a, 0, i16p), P(dq, 0, i32p), qp); o.xo_dequant_4x4_dc(X._ptr(b2, X.i16p), X._ptr(dq, X.i32p), qp)
        assert np.array_equal(a, b2)
        for fn, n in ((d.add8x8_idct_dc, 4), (d.add16x16_idct_dc, 16)):
            r1, r2 = fd.copy(), fd.copy()
            fn(P(r1), P(a, 0, i16p)); o.xo_add_idct_dc(X._ptr(r2), X._ptr(a, X.i16p), n)
            assert np.array_equal(r1, r2)


def test_mc_table(pkg, L, port):
    m = McTable()
    assert L.x264_mc_init_cuda(C.byref(m)) == 0
    from x264_vs2008_b200 import synth
    w, h = 96, 64
    g = port.geometry(w, h)
    plane = port.plane_from_picture(g, synth.Clip(w, h, seed=2).luma(0))
    fh, fv, fc, _ = port.frame_filter(g, plane, 0, want_integral=False)
    # hpel_filter on a 48x10 window (checkasm.c:822-850 compares bytes 2..44 of each row; we compare all the C body defines)
    dh, dv, dc = np.zeros_like(plane), np.zeros_like(plane), np.zeros_like(plane)
    o0 = g.origin + 8 * g.stride + 8
    m.hpel_filter(P(dh, o0), P(dv, o0), P(dc, o0), P(plane, o0), g.stride, 48, 10, None)
    for y in range(10):
        r = o0 + y * g.stride
        assert np.array_equal(dh[r:r + 48], fh[r:r + 48]) and np.array_equal(dc[r:r + 48], fc[r:r + 48])
        assert np.array_equal(dv[r - 2:r + 51], fv[r - 2:r + 51])
    # frame_init_lowres_core w=40 (checkasm.c:852-881)
    outs = [np.zeros(64 * 24, np.uint8) for _ in range(4)]
    m.frame_init_lowres_core(P(plane, g.origin), P(outs[0]), P(outs[1]), P(outs[2]), P(outs[3]), g.stride, 64, 40, 20)
    want = port.init_lowres(g, plane.copy())
    for k in range(4):
        for y in range(20):
            assert np.array_equal(outs[k][y * 64:y * 64 + 40], want[k][g.origin_lowres + y * g.stride_lowres:][:40])
    # mc_chroma (checkasm.c:748-768) and avg with plain / implicit weights (:770-800)
    rngc = np.random.default_rng(14)
    src = rngc.integers(0, 256, (64, 64), dtype=np.uint8)
    for (bw, bh) in ((8, 8), (8, 4), (4, 8), (4, 4), (4, 2), (2, 4), (2, 2)):
        for _ in range(6):
            mvx, mvy = int(rngc.integers(-60, 61)), int(rngc.integers(-60, 61))
            d1, d2 = np.full((8, 16), 9, np.uint8), np.full((8, 16), 9, np.uint8)
            m.mc_chroma(P(d1), 16, P(src, 24 * 64 + 24), 64, mvx, mvy, bw, bh)
            port.lib.xo_mc_chroma(X._ptr(d2), 16, X._ptr(src, X.u8p, 24 * 64 + 24), 64, mvx, mvy, bw, bh)
            assert np.array_equal(d1[:bh, :bw], d2[:bh, :bw]) and (d1[:, bw + 2:] == 9).all(), ("mc_chroma", bw, bh, mvx, mvy)
    s1, s2 = rngc.integers(0, 256, (16, 32), dtype=np.uint8), rngc.integers(0, 256, (16, 48), dtype=np.uint8)
    s1[0, :4], s2[0, :4] = (0, 255, 0, 255), (255, 0, 255, 0)
    for ip, (bw, bh) in enumerate(((16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (4, 2), (2, 4), (2, 2))):
        for weight in (32, 21, 43, 5, 59, -10, 74):
            d1, d2 = np.full((16, 16), 7, np.uint8), np.full((16, 16), 7, np.uint8)
            m.avg[ip](P(d1), 16, P(s1), 32, P(s2), 48, weight)
            port.lib.xo_pixel_avg(ip, X._ptr(d2), 16, X._ptr(s1), 32, X._ptr(s2), 48, weight)
            assert np.array_equal(d1, d2), ("avg", ip, weight)
    # mc_luma / get_ref at random qpel vectors (checkasm.c:702-746)
    rng = np.random.default_rng(4)
    planes = [plane, fh, fv, fc]
    arr = (u8p * 4)(*[P(p, g.origin + 20 * g.stride + 30) for p in planes])
    arr_o = (X.u8p * 4)(*[X._ptr(p, X.u8p, g.origin + 20 * g.stride + 30) for p in planes])
    for (bw, bh) in ((16, 16), (16, 8), (8, 16), (8, 8), (8, 4), (4, 8), (4, 4), (20, 18), (12, 10)):
        for _ in range(8):
            mvx, mvy = int(rng.integers(-40, 41)), int(rng.integers(-40, 41))
            d1, d2 = np.zeros(32 * 20, np.uint8), np.zeros(32 * 20, np.uint8)
            m.mc_luma(P(d1), 32, arr, g.stride, mvx, mvy, bw, bh)
            port.lib.xo_mc_luma(X._ptr(d2), 32, arr_o, g.stride, mvx, mvy, bw, bh)
            assert np.array_equal(d1, d2), (bw, bh, mvx, mvy)
            st = C.c_int(32)
            d3 = np.zeros(32 * 20, np.uint8)
            res = m.get_ref(P(d3), C.byref(st), arr, g.stride, mvx, mvy, bw, bh)
            got = np.ctypeslib.as_array(res, shape=(20 * st.value,))
            for y in range(bh):
                assert np.array_equal(got[y * st.value:y * st.value + bw], d2[y * 32:y * 32 + bw])
    L.x264_cuda_tables_shutdown()
