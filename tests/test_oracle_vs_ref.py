"""not-gpu: oracle port vs the unmodified reference, live (only where oracle/_ref is built), on fresh seeded inputs —
the checkasm-style differential test (S/tools/checkasm.c) with the reference as the other side."""
import ctypes as C

import numpy as np
import pytest
import xo_api as X


def test_cost_tables_all_qps(port, ref):
    for q in range(52):
        assert np.array_equal(port.cost_mv_table(q), ref.cost_mv_table(q)), q


def test_pixel_metrics_random(port, ref):
    rng = np.random.default_rng(7)
    a = rng.integers(0, 256, (64, 64), dtype=np.uint8)
    b = rng.integers(0, 256, (64, 64), dtype=np.uint8)
    for m in range(4):
        for ip in range(7):
            if m == X.SA8D and ip not in (0, 3):
                continue
            for _ in range(64):  # 64 misalignments like checkasm.c:227
                oa, ob = int(rng.integers(0, 32)) * 64 + int(rng.integers(0, 32)), int(rng.integers(0, 32)) * 64 + int(rng.integers(0, 32))
                assert port.pixel_cmp(m, ip, a, 64, b, 64, oa, ob) == ref.pixel_cmp(m, ip, a, 64, b, 64, oa, ob)
    for ip in (0, 3):
        assert port.lib.xo_pixel_var(ip, X._ptr(a), 64) == ref.lib.xo_pixel_var(ip, X._ptr(a), 64)
    for ip in range(4):
        assert port.lib.xo_pixel_hadamard_ac(ip, X._ptr(a), 64) == ref.lib.xo_pixel_hadamard_ac(ip, X._ptr(a), 64)


def test_ads(port, ref):
    """checkasm.c:433-464: 100 trials, sums[72], dc 14-bit, width 28, delta 32"""
    rng = np.random.default_rng(9)
    for ip in range(7):
        for _ in range(100):
            sums = rng.integers(0, 1 << 14, 72, dtype=np.uint16)
            dc = (C.c_int * 4)(*[int(x) for x in rng.integers(0, 1 << 14, 4)])
            cost = rng.integers(0, 1 << 10, 32, dtype=np.uint16)
            thresh = int(rng.integers(0, 1 << 15))
            m1, m2 = np.zeros(32, np.int16), np.zeros(32, np.int16)
            n1 = port.lib.xo_pixel_ads(ip, dc, X._ptr(sums, X.u16p), 32, X._ptr(cost, X.u16p), X._ptr(m1, X.i16p), 28, thresh)
            n2 = ref.lib.xo_pixel_ads(ip, dc, X._ptr(sums, X.u16p), 32, X._ptr(cost, X.u16p), X._ptr(m2, X.i16p), 28, thresh)
            assert n1 == n2 and np.array_equal(m1[:n1], m2[:n2])


def test_frame_ops(pkg, port, ref):
    from x264_vs2008_b200 import synth
    for (w, h) in ((64, 48), (100, 70), (352, 288)):
        g = port.geometry(w, h)
        gr = ref.geometry(w, h)
        assert all(getattr(g, f[0]) == getattr(gr, f[0]) for f in g._fields_)
        pic = synth.Clip(w, h, seed=w).luma(0)
        pp, pr = port.plane_from_picture(g, pic), ref.plane_from_picture(g, pic)
        assert np.array_equal(pp, pr)
        for s8 in (0, 1):
            for x, y in zip(port.frame_filter(g, pp, s8), ref.frame_filter(g, pr, s8)):
                assert np.array_equal(x, y)
        la, lb = port.init_lowres(g, pp.copy()), ref.init_lowres(g, pr.copy())
        for x, y in zip(la, lb):
            x2, y2 = x.reshape(-1, g.stride_lowres), y.reshape(-1, g.stride_lowres)
            # odd mb_width: columns >= width_lowres (+ right border) are never written by the reference
            # (stale memory of a recycled frame buffer) -> only the defined area is comparable
            c1 = g.stride_lowres if g.mb_width % 2 == 0 else X.PADH + g.width_lowres
            assert np.array_equal(x2[:, :c1], y2[:, :c1])


def test_dct_quant(port, ref):
    rng = np.random.default_rng(11)
    for trial in range(200):
        p1 = rng.integers(0, 256, 16 * 16, dtype=np.uint8)
        p2 = rng.integers(0, 256, 32 * 16, dtype=np.uint8)
        if trial % 10 == 0:
            p1[:], p2[:] = 255 * (trial % 20 == 0), 255 * (trial % 20 != 0)
        for fn, n in (("xo_sub4x4_dct", 16), ("xo_sub8x8_dct8", 64)):
            d1, d2 = np.zeros(n, np.int16), np.zeros(n, np.int16)
            getattr(port.lib, fn)(X._ptr(d1, X.i16p), X._ptr(p1), X._ptr(p2))
            getattr(ref.lib, fn)(X._ptr(d2, X.i16p), X._ptr(p1), X._ptr(p2))
            assert np.array_equal(d1, d2)
            for cqm in (0, 1):
                qp = int(rng.integers(0, 52))
                lst = int(rng.integers(0, 4 if n == 16 else 2))
                mf1, b1, mf2, b2 = (np.zeros(n, np.uint16) for _ in range(4))
                tfn = "xo_quant4_tables" if n == 16 else "xo_quant8_tables"
                getattr(port.lib, tfn)(cqm, lst, qp, X._ptr(mf1, X.u16p), X._ptr(b1, X.u16p))
                getattr(ref.lib, tfn)(cqm, lst, qp, X._ptr(mf2, X.u16p), X._ptr(b2, X.u16p))
                assert np.array_equal(mf1, mf2) and np.array_equal(b1, b2), (cqm, lst, qp)
                dq1, dq2 = np.zeros(6 * n, np.int32), np.zeros(6 * n, np.int32)
                dfn = "xo_dequant4_table" if n == 16 else "xo_dequant8_table"
                getattr(port.lib, dfn)(cqm, lst, X._ptr(dq1, X.i32p))
                getattr(ref.lib, dfn)(cqm, lst, X._ptr(dq2, X.i32p))
                assert np.array_equal(dq1, dq2)
                q1, q2 = d1.copy(), d1.copy()
                qfn = "xo_quant_4x4" if n == 16 else "xo_quant_8x8"
                nz1 = getattr(port.lib, qfn)(X._ptr(q1, X.i16p), X._ptr(mf1, X.u16p), X._ptr(b1, X.u16p))
                nz2 = getattr(ref.lib, qfn)(X._ptr(q2, X.i16p), X._ptr(mf1, X.u16p), X._ptr(b1, X.u16p))
                assert nz1 == nz2 and np.array_equal(q1, q2)
                dfn2 = "xo_dequant_4x4" if n == 16 else "xo_dequant_8x8"
                getattr(port.lib, dfn2)(X._ptr(q1, X.i16p), X._ptr(dq1, X.i32p), qp)
                getattr(ref.lib, dfn2)(X._ptr(q2, X.i16p), X._ptr(dq1, X.i32p), qp)
                assert np.array_equal(q1, q2)
                r1, r2 = p2.copy(), p2.copy()
                ifn = "xo_add4x4_idct" if n == 16 else "xo_add8x8_idct8"
                getattr(port.lib, ifn)(X._ptr(r1), X._ptr(q1, X.i16p))
                getattr(ref.lib, ifn)(X._ptr(r2), X._ptr(q2, X.i16p))
                assert np.array_equal(r1, r2) and np.array_equal(q1, q2)
        # DC paths: +-4080 extremes and 13-bit random (checkasm.c:577-589)
        dc = rng.integers(-4096, 4096, 16).astype(np.int16) if trial % 3 else np.full(16, 4080 * (1 - 2 * (trial % 2)), np.int16)
        a, b = dc.copy(), dc.copy()
        port.lib.xo_dct4x4dc(X._ptr(a, X.i16p)); ref.lib.xo_dct4x4dc(X._ptr(b, X.i16p))
        assert np.array_equal(a, b)
        qp = int(rng.integers(0, 52))
        mf, bias = int(rng.integers(100, 30000)), int(rng.integers(0, 30000))
        assert port.lib.xo_quant_4x4_dc(X._ptr(a, X.i16p), mf, bias) == ref.lib.xo_quant_4x4_dc(X._ptr(b, X.i16p), mf, bias)
        assert np.array_equal(a, b)
        a2, b2 = a[:4].copy(), a[:4].copy()
        assert port.lib.xo_quant_2x2_dc(X._ptr(a2, X.i16p), mf, bias) == ref.lib.xo_quant_2x2_dc(X._ptr(b2, X.i16p), mf, bias)
        assert np.array_equal(a2, b2)
        port.lib.xo_idct4x4dc(X._ptr(a, X.i16p)); ref.lib.xo_idct4x4dc(X._ptr(b, X.i16p))
        assert np.array_equal(a, b)
        dq = np.zeros(6 * 16, np.int32)
        port.lib.xo_dequant4_table(0, 0, X._ptr(dq, X.i32p))
        port.lib.xo_dequant_4x4_dc(X._ptr(a, X.i16p), X._ptr(dq, X.i32p), qp); ref.lib.xo_dequant_4x4_dc(X._ptr(b, X.i16p), X._ptr(dq, X.i32p), qp)
        assert np.array_equal(a, b)
        for n in (4, 16):
            r1, r2 = p2.copy(), p2.copy()
            port.lib.xo_add_idct_dc(X._ptr(r1), X._ptr(a, X.i16p), n); ref.lib.xo_add_idct_dc(X._ptr(r2), X._ptr(a, X.i16p), n)
            assert np.array_equal(r1, r2)


@pytest.mark.parametrize("size,method,satd,weighted", [((176, 144), X.ME_HEX, 1, 0), ((176, 144), X.ME_DIA, 0, 0), ((208, 112), X.ME_HEX, 1, 1),
                                                       ((32, 64), X.ME_HEX, 1, 0)])
def test_lowres_lookahead(pkg, port, ref, size, method, satd, weighted):
    """x264_slicetype_frame_cost through the reference's x264_rc_analyse_slice vs the port, whole evaluation schedule"""
    from x264_vs2008_b200 import synth
    from helpers import lowres_planes, oracle_lookahead, lookahead_digest
    w, h = size
    g = port.geometry(w, h)
    planes = lowres_planes(ref, g, synth.Clip(w, h, seed=31), 3)
    a = oracle_lookahead(port, g, planes, method, 16, satd, weighted)
    b = oracle_lookahead(ref, g, planes, method, 16, satd, weighted, is_ref=True)
    sa, xa = lookahead_digest(a, g)
    sb, xb = lookahead_digest(b, g)
    assert np.array_equal(sa, sb), (sa, sb)
    assert np.array_equal(xa, xb)


@pytest.mark.parametrize("size,aq,vbv", [((176, 144), True, True), ((208, 112), False, True), ((32, 64), True, True), ((176, 144), True, False)])
def test_lowres_lookahead_vbv(pkg, port, ref, size, aq, vbv):
    """the VBV form of x264_slicetype_frame_cost (slicetype.c:300-316): every block evaluated, per-row sums, AQ-weighted costs"""
    from x264_vs2008_b200 import synth
    from helpers import lowres_planes, oracle_lookahead, lookahead_digest
    w, h = size
    g = port.geometry(w, h)
    planes = lowres_planes(ref, g, synth.Clip(w, h, seed=33), 3)
    rng = np.random.default_rng(5)
    inv = [rng.integers(128, 512, g.mb_width * g.mb_height).astype(np.uint16) for _ in range(3)] if aq else None
    a = oracle_lookahead(port, g, planes, X.ME_HEX, 16, 1, 0, vbv=vbv, inv_qscale=inv)
    b = oracle_lookahead(ref, g, planes, X.ME_HEX, 16, 1, 0, is_ref=True, vbv=vbv, inv_qscale=inv)
    sa, xa = lookahead_digest(a, g, all_blocks=vbv)
    sb, xb = lookahead_digest(b, g, all_blocks=vbv)
    assert np.array_equal(sa, sb), (sa, sb)
    assert np.array_equal(xa, xb)
    small = g.mb_width <= 2 or g.mb_height <= 2
    for ra, rb in zip(a, b):
        if vbv and not small:  # tiny frames take the first branch of slicetype.c:293-298: no row sums
            assert np.array_equal(ra[9], rb[9]), (ra[0], ra[9], rb[9])
        assert ra[10] == rb[10], (ra[0], ra[10], rb[10])


@pytest.mark.parametrize("size,kw", [((176, 144), dict()), ((176, 144), dict(slice_b=1)), ((208, 112), dict(cavlc_8x8dct=1, alpha=-2, beta=2, chroma_off=3)),
                                     ((64, 48), dict(chaos=True)), ((96, 80), dict(chaos=True, slice_b=1, cavlc_8x8dct=1, alpha=6, beta=-4, chroma_off=-5)),
                                     ((176, 144), dict(psub8x8=0, qp_centre=18, alpha=-6, beta=-6)), ((32, 16), dict(qp_centre=45, alpha=12, beta=12))])
def test_deblock(pkg, port, ref, size, kw):
    """x264_frame_deblock_row over a whole frame: reference vs port, encoder-like and fully random macroblock state"""
    from x264_vs2008_b200 import synth
    from helpers import make_deblock_info, blocky_recon
    w, h = size
    g = port.geometry(w, h)
    for seed in range(3):
        info = make_deblock_info(g, seed=100 + seed, **kw)
        y, u, v = blocky_recon(synth.Clip(w, h, seed=5), g, seed=seed)
        outs = []
        for o in (port, ref):
            py = o.new_plane(g)
            pad = py.reshape(-1, g.stride)
            pad[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]] = y
            uu, vv = u.copy(), v.copy()
            o.frame_deblock(g, info, py, uu, vv)
            outs.append((py.copy(), uu, vv))
        assert not np.array_equal(outs[0][0].reshape(-1, g.stride)[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]], y)  # something was filtered
        for a, b in zip(*outs):
            assert np.array_equal(a, b)


def test_chroma_me(pkg, port, ref):
    """refine_subpel with b_chroma_me (P slices, subme >= 5): COST_MV_SATD adds mc_chroma + mbcmp[i_pixel+3] of U and V (me.c:655-677)"""
    from x264_vs2008_b200 import synth
    from helpers import make_me_jobs, padded_chroma
    w, h = 160, 128
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=21)
    (y1, u1, v1), (y0, u0, v0) = clip.yuv420(1), clip.yuv420(0)
    pe, pr = port.plane_from_picture(g, y1), port.plane_from_picture(g, y0)
    fh, fv, fc, integ = port.frame_filter(g, pr, 1)
    chroma = [padded_chroma(g, c) for c in (u1, v1, u0, v0)]
    n_diff = 0
    for method in (X.ME_DIA, X.ME_HEX, X.ME_ESA):
        _, mis = make_me_jobs(pkg, g, seed=300 + method, n=60, me_range=16, qp=(12, 26, 40), pixels=(0, 1, 2, 3))
        for subme in (5, 6, 7):
            for mi in mis:
                mi.me_method = method
                mi.bx, mi.by = (mi.bx // 16) * 16, (mi.by // 16) * 16
                a = port.me_search_subpel_chroma(g, pe, [pr, fh, fv, fc], integ, chroma, mi, subme, 1)
                b = ref.me_search_subpel_chroma(g, pe, [pr, fh, fv, fc], integ, chroma, mi, subme, 1)
                assert (a.mv[0], a.mv[1], a.cost, a.cost_mv) == (b.mv[0], b.mv[1], b.cost, b.cost_mv)
                c = port.me_search_subpel(g, pe, [pr, fh, fv, fc], integ, mi, subme, 1)
                n_diff += (c.mv[0], c.mv[1]) != (a.mv[0], a.mv[1]) or c.cost != a.cost
    assert n_diff > 50  # the chroma term really changes costs/decisions on this content


def _metric_inputs(port, w, h, seed):
    from x264_vs2008_b200 import synth
    from helpers import blocky_recon
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=seed)
    y, u, v = clip.yuv420(1)
    ry, ru, rv = blocky_recon(clip, g, frame=1, seed=seed)  # a "reconstruction" of the same picture
    H16, W16 = 16 * g.mb_height, 16 * g.mb_width
    def pad(a, hh, ww):
        out = np.zeros((hh, ww), np.uint8)
        out[:a.shape[0], :a.shape[1]] = a
        return out
    return g, (pad(y, H16, W16), pad(u, H16 // 2, W16 // 2), pad(v, H16 // 2, W16 // 2)), (ry, ru, rv)


@pytest.mark.parametrize("size", [(176, 144), (100, 70), (352, 288), (20, 12)])
def test_frame_metrics(pkg, port, ref, size):
    """x264_pixel_ssd_wxh, x264_pixel_ssim_wxh, ac_energy_mb / x264_adaptive_quant_frame, hadamard_ac: reference vs port, and the
    product's host-side float tails (x264_cuda_host_ssim_end, x264_cuda_host_aq) vs the reference"""
    w, h = size
    g, (y, u, v), (ry, ru, rv) = _metric_inputs(port, w, h, seed=9)
    for (a, b, ww, hh) in ((y, ry, w, h), (u, ru, w // 2, h // 2), (v, rv, w // 2, h // 2), (y, ry, w - 3, h - 5)):
        assert port.frame_ssd(a, b, ww, hh) == ref.frame_ssd(a, b, ww, hh) > 0
        if ww >= 8 and hh >= 8:
            s_ref = ref.frame_ssim(a, b, ww, hh)
            assert port.frame_ssim(a, b, ww, hh) == s_ref
            sums = port.frame_ssim_sums(a, b, ww, hh)
            assert np.array_equal(sums, ref.frame_ssim_sums(a, b, ww, hh))
            assert float(pkg.lib().x264_cuda_host_ssim_end(sums.ctypes.data, ww // 4, hh // 4)) == s_ref
    py = port.new_plane(g)
    py.reshape(-1, g.stride)[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]] = y
    e = port.frame_mb_energy(g, py, u, v)
    assert np.array_equal(e, ref.frame_mb_energy(g, py, u, v))
    assert np.array_equal(port.frame_mb_hadamard_ac(g, py), ref.frame_mb_hadamard_ac(g, py))
    for strength in (1.0, 0.6, 2.5):
        q_ref, i_ref = ref.frame_aq(g, py, u, v, strength)
        q_port, i_port = port.frame_aq(g, py, u, v, strength)
        q_host, i_host = pkg.host_aq(e, strength)
        assert np.array_equal(q_ref, q_port) and np.array_equal(i_ref, i_port)
        assert np.array_equal(q_ref, q_host) and np.array_equal(i_ref, i_host)


def test_aq_tables_all_energies(pkg, port, ref):
    """every leading-zero count and every 7-bit mantissa of the log2 table, through x264_adaptive_quant_frame on crafted pictures
    is impractical; instead sweep the host helper against the port over a dense set of energies (both rebuild the tables from
    log2/exp2), and pin the port to the reference on real pictures above"""
    e = np.unique(np.concatenate([np.arange(1, 70000, 7), (np.arange(128, 256)[:, None] << np.arange(0, 24)[None, :]).ravel(),
                                  np.array([1, 2, 3, 0x7fffffff, 0xffffffff])]).astype(np.uint64)).astype(np.uint32)
    for strength in (1.0, 3.0):
        q, i = pkg.host_aq(e, strength)
        q2, i2 = np.zeros(len(e), np.float32), np.zeros(len(e), np.uint16)
        port.lib.xo_aq_from_energy(e.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_uint32)), len(e), __import__("ctypes").c_float(strength),
                                   q2.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float)), X._ptr(i2, X.u16p))
        assert np.array_equal(q, q2) and np.array_equal(i, i2)


def test_pixel_avg(port, ref):
    """h->mc.avg[10] (mc.c:52-125): rounded average and implicit weighted bi-prediction, incl. weights that clip"""
    rng = np.random.default_rng(12)
    for ip in range(10):
        for weight in (32, 21, 43, 0, 64, 11, 53, -20, 84):
            a = rng.integers(0, 256, (16, 16), dtype=np.uint8)
            b = rng.integers(0, 256, (16, 16), dtype=np.uint8)
            if weight in (-20, 84):
                a[:4], b[:4] = 255, 0
            d1, d2 = np.zeros((16, 16), np.uint8), np.zeros((16, 16), np.uint8)
            port.lib.xo_pixel_avg(ip, X._ptr(d1), 16, X._ptr(a), 16, X._ptr(b), 16, weight)
            ref.lib.xo_pixel_avg(ip, X._ptr(d2), 16, X._ptr(a), 16, X._ptr(b), 16, weight)
            assert np.array_equal(d1, d2), (ip, weight)


@pytest.mark.parametrize("me_range", [16, 24, 8])
def test_umh(pkg, port, ref, me_range):
    """--me umh (me.c:306-447): uneven cross, early terminations, adaptive range from predictor agreement, hexagon grid, hex2 refine"""
    from x264_vs2008_b200 import synth
    from helpers import make_me_jobs
    w, h = 160, 128
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=21)
    pe, pr = port.plane_from_picture(g, clip.luma(1)), port.plane_from_picture(g, clip.luma(0))
    fh, fv, fc, integ = port.frame_filter(g, pr, 1)
    for spread, centre in ((48, None), (10, (-20, -12)), (4, (-20, -12))):
        _, mis = make_me_jobs(pkg, g, seed=500 + me_range + spread, n=150, me_range=me_range, qp=(12, 26, 40), pixels=(0, 1, 2, 3, 4, 5, 6), mvp_spread=spread,
                              centre=centre)
        for i, mi in enumerate(mis):
            mi.me_method = X.ME_UMH
            mi.b_sub8x8 = 1
            if centre is not None and i % 2:  # neighbours that agree with the predictor: exercises the small-range contexts
                for k in range(mi.i_mvc):
                    mi.mvc[k][0], mi.mvc[k][1] = mi.mvp[0] + (k % 3) - 1, mi.mvp[1] + (k % 2)
            for subme in (1, 2, 5):
                a = port.me_search_subpel(g, pe, [pr, fh, fv, fc], integ, mi, subme, 1)
                b = ref.me_search_subpel(g, pe, [pr, fh, fv, fc], integ, mi, subme, 1)
                assert (a.mv[0], a.mv[1], a.cost, a.cost_mv) == (b.mv[0], b.mv[1], b.cost, b.cost_mv), (spread, i, subme, mi.i_pixel)


def test_refine_qpel(pkg, port, ref):
    """x264_me_refine_qpel (me.c:633-643): refine_subpel with b_refine_qpel = 1 from a given vector, with and without chroma ME"""
    from x264_vs2008_b200 import synth
    from helpers import make_me_jobs, padded_chroma
    w, h = 160, 128
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=21)
    (y1, u1, v1), (y0, u0, v0) = clip.yuv420(1), clip.yuv420(0)
    pe, pr = port.plane_from_picture(g, y1), port.plane_from_picture(g, y0)
    fh, fv, fc, _ = port.frame_filter(g, pr, 0, want_integral=False)
    chroma = [padded_chroma(g, c) for c in (u1, v1, u0, v0)]
    rng = np.random.default_rng(77)
    _, mis = make_me_jobs(pkg, g, seed=321, n=120, me_range=16, qp=(12, 26, 40), pixels=(0, 1, 2, 3, 4, 5, 6), mvp_spread=30, centre=(-20, -12))
    for i, mi in enumerate(mis):
        if mi.i_pixel <= 3:
            mi.bx, mi.by = (mi.bx // 8) * 8, (mi.by // 8) * 8
        mv = [int(rng.integers(-12, 13)) - 20, int(rng.integers(-12, 13)) - 12]
        cost = int(rng.integers(200, 6000))
        for subme in (1, 2, 3, 5, 7):
            for ch in (None, chroma):
                a = port.me_refine_qpel(g, pe, [pr, fh, fv, fc], ch, mi, subme, 1, mv, cost)
                b = ref.me_refine_qpel(g, pe, [pr, fh, fv, fc], ch, mi, subme, 1, mv, cost)
                assert (a.mv[0], a.mv[1], a.cost, a.cost_mv) == (b.mv[0], b.mv[1], b.cost, b.cost_mv), (i, subme, ch is not None, mi.i_pixel)


def bidir_cases(pkg, g, seed, n):
    """seeded (MeIn, mvp0, mvp1, mv0, mv1, weight) for the bidirectional refinement"""
    from helpers import make_me_jobs
    rng = np.random.default_rng(seed)
    jobs, mis = make_me_jobs(pkg, g, seed=seed, n=n, me_range=16, qp=(12, 26, 40), pixels=(0, 1, 2, 3), mvp_spread=8, centre=(-20, -12))
    out = []
    for j, mi in zip(jobs, mis):
        mi.bx, mi.by = (mi.bx // 8) * 8, (mi.by // 8) * 8
        j["bx"], j["by"] = mi.bx, mi.by
        mvp0 = [int(rng.integers(-10, 11)) - 20, int(rng.integers(-10, 11)) - 12]
        mvp1 = [int(rng.integers(-10, 11)) + 20, int(rng.integers(-10, 11)) + 12]
        mv0 = [mvp0[0] + int(rng.integers(-6, 7)), mvp0[1] + int(rng.integers(-6, 7))]
        mv1 = [mvp1[0] + int(rng.integers(-6, 7)), mvp1[1] + int(rng.integers(-6, 7))]
        if rng.integers(0, 25) == 0:
            mv0[1] = mi.mv_max_spel[1] - int(rng.integers(0, 8))  # the early return of me.c:874-876
        out.append((j, mi, mvp0, mvp1, mv0, mv1, int(rng.choice([32, 32, 21, 43, 27]))))
    return out


def test_refine_bidir_satd(pkg, port, ref):
    """x264_me_refine_bidir_satd (me.c:843-927): 32 candidate pairs per pass in the reference's order, the aliasing visited map,
    the quirk of clipping mvp[1] with the x limits"""
    from x264_vs2008_b200 import synth
    w, h = 160, 128
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=21)
    pe = port.plane_from_picture(g, clip.luma(1))
    refs = []
    for fr in (0, 2):
        pr = port.plane_from_picture(g, clip.luma(fr))
        fh, fv, fc, _ = port.frame_filter(g, pr, 0, want_integral=False)
        refs.append([pr, fh, fv, fc])
    moved = 0
    for j, mi, mvp0, mvp1, mv0, mv1, weight in bidir_cases(pkg, g, 55, 150):
        for satd in (1, 0):
            a0, a1, _ = port.me_refine_bidir_satd(g, pe, refs[0], refs[1], mi, mvp0, mvp1, weight, satd, mv0, mv1)
            b0, b1, _ = ref.me_refine_bidir_satd(g, pe, refs[0], refs[1], mi, mvp0, mvp1, weight, satd, mv0, mv1)
            assert (a0, a1) == (b0, b1), (mi.i_pixel, mvp0, mvp1, mv0, mv1, weight, satd)
            moved += (a0, a1) != (tuple(mv0), tuple(mv1))
    assert moved > 100


@pytest.mark.parametrize("cqm", [0, 1])
def test_residual_inter_mb(port, ref, cqm):
    """the inter branch of x264_macroblock_encode (+ chroma) on hand-loaded macroblocks: port vs the reference's own function"""
    import helpers
    for i, (qp, cqp, fy, fu, fv, py, pu, pv) in enumerate(helpers.skip_probe_cases(50 + cqm, 240)):
        for flags in range(4):
            rin = X.ResidIn(qp, min(cqp, 51), flags & 1, flags >> 1, cqm)
            a, b = port.residual_inter_mb(rin, fy, fu, fv, py, pu, pv), ref.residual_inter_mb(rin, fy, fu, fv, py, pu, pv)
            tag = (i, qp, cqp, flags)
            assert (a[0].cbp_luma, a[0].cbp_chroma) == (b[0].cbp_luma, b[0].cbp_chroma), tag
            assert bytes(a[0].nnz) == bytes(b[0].nnz), tag
            if flags & 1:
                assert np.array_equal(np.array(a[0].luma8x8), np.array(b[0].luma8x8)), tag
                assert np.array_equal(np.array(a[0].luma4x4)[16:], np.array(b[0].luma4x4)[16:]), tag
            else:
                assert np.array_equal(np.array(a[0].luma4x4), np.array(b[0].luma4x4)), tag
            assert np.array_equal(np.array(a[0].chroma_dc), np.array(b[0].chroma_dc)), tag
            for k in (1, 2, 3):
                assert np.array_equal(a[k], b[k]), tag


def intra16_neighbours(rng, fy, fu, fv, flat):
    """neighbour pixels (corner, row above, left column) resembling the source block (so that the residual is small) or random"""
    out = []
    for t, n in ((fy, 16), (fu, 8), (fv, 8)):
        t2 = t.reshape(n, n).astype(np.int32)
        nb = np.concatenate([[t2[0, 0]], t2[0], t2[:, 0]]) + rng.integers(-6, 7, 2 * n + 1)
        if not flat:
            nb = rng.integers(0, 256, 2 * n + 1)
        out.append(np.clip(nb, 0, 255).astype(np.uint8))
    return out


@pytest.mark.parametrize("cqm", [0, 1])
def test_residual_intra16_mb(port, ref, cqm):
    """I_16x16 macroblocks through x264_macroblock_encode (predict + x264_mb_encode_i16x16 + intra chroma): port vs the reference"""
    import helpers
    rng = np.random.default_rng(70 + cqm)
    seen_dc_only = seen_decimated = 0
    for i, (qp, cqp, fy, fu, fv, py, pu, pv) in enumerate(helpers.skip_probe_cases(70 + cqm, 300)):
        nby, nbu, nbv = intra16_neighbours(rng, fy, fu, fv, flat=i % 3 != 0)
        mode16, modec, dec = int(rng.integers(0, 7)), int(rng.integers(0, 7)), int(rng.integers(0, 2))
        rin = X.ResidIn(qp, min(cqp, 51), 0, dec, cqm)
        a = port.residual_intra16_mb(rin, mode16, modec, fy, fu, fv, nby, nbu, nbv)
        b = ref.residual_intra16_mb(rin, mode16, modec, fy, fu, fv, nby, nbu, nbv)
        tag = (i, qp, cqp, mode16, modec, dec)
        assert (a[0].cbp_luma, a[0].cbp_chroma) == (b[0].cbp_luma, b[0].cbp_chroma), tag
        assert bytes(a[0].nnz) == bytes(b[0].nnz), tag
        assert np.array_equal(np.array(a[0].luma4x4), np.array(b[0].luma4x4)), tag
        assert np.array_equal(np.array(a[0].chroma_dc), np.array(b[0].chroma_dc)), tag
        assert np.array_equal(a[1], b[1]), tag
        for k in (2, 3, 4):
            assert np.array_equal(a[k], b[k]), tag
        seen_dc_only += a[0].cbp_luma == 0 and a[0].nnz[24] != 0
        seen_decimated += dec and a[0].cbp_luma == 0
    assert seen_dc_only and seen_decimated  # both reconstruction branches of macroblock.c:264-269 were exercised


@pytest.mark.parametrize("cqm", [0, 1])
def test_probe_skip(port, ref, cqm):
    """x264_macroblock_probe_skip (b_bidir form) + the lambda2 table its chroma gate reads"""
    import helpers
    for q in range(52):
        assert port.lib.xo_lambda2(q) == ref.lib.xo_lambda2(q), q
    n1 = 0
    cases = helpers.skip_probe_cases(60 + cqm, 1500)
    for i, (qp, cqp, fy, fu, fv, py, pu, pv) in enumerate(cases):
        rin = X.ResidIn(qp, cqp, 0, 1, cqm)
        a, b = port.probe_skip_mb(rin, fy, fu, fv, py, pu, pv), ref.probe_skip_mb(rin, fy, fu, fv, py, pu, pv)
        assert a == b, (i, qp, cqp)
        n1 += a
    assert len(cases) // 5 < n1 < 4 * len(cases) // 5, n1  # both outcomes well represented


def test_intra_predictors_and_mb_costs(port, ref):
    """predict_16x16 / predict_8x8c (all seven modes each) and the Intra16x16 + chroma cost stage of x264_mb_analyse_intra(_chroma)"""
    import helpers
    modes16, modesc = set(), set()
    for i, (nbr, lam, satd, sb, fy, fu, fv, nby, nbu, nbv) in enumerate(helpers.intra_cases(90, 600)):
        for m in range(7):
            assert np.array_equal(port.predict(0, m, nby), ref.predict(0, m, nby)), (i, "16x16", m)
            assert np.array_equal(port.predict(1, m, nbu), ref.predict(1, m, nbu)), (i, "8x8c", m)
        iin = X.IntraIn(nbr, lam, satd, sb)
        a, b = port.intra_mb_costs(iin, fy, fu, fv, nby, nbu, nbv), ref.intra_mb_costs(iin, fy, fu, fv, nby, nbu, nbv)
        assert a.astuple() == b.astuple(), (i, nbr, lam, satd, sb)
        modes16.add(a.mode16); modesc.add(a.mode_chroma)
    assert modes16 == set(range(7)) and modesc == set(range(7)), (modes16, modesc)  # every mode wins somewhere
