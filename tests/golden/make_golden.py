#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_harness.so, i.e. the x264
snapshot's own C compiled in place).  Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference ships no golden vectors, so these are its outputs on seeded inputs; the
not-gpu tests require the oracle port to reproduce every byte, the gpu tests require the CUDA path to."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import xo_api as X  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

pkg = ge.load_pkg()
from x264_vs2008_b200 import synth  # noqa: E402
from helpers import make_me_jobs, lowres_planes, oracle_lookahead, lookahead_digest, make_deblock_info, blocky_recon, skip_probe_cases, residual_digest, intra_cases, intra_digest  # noqa: E402


LOOKAHEAD_GOLDEN = (("qcif_hex_satd", (176, 144), X.ME_HEX, 1, 0), ("qcif_dia_sad", (176, 144), X.ME_DIA, 0, 0),
                    ("w208_hex_satd_wb", (208, 112), X.ME_HEX, 1, 1), ("tiny", (32, 64), X.ME_HEX, 1, 0))


DEBLOCK_GOLDEN = (("qcif_p", (176, 144), dict()), ("qcif_b", (176, 144), dict(slice_b=1)),
                  ("w208_cavlc8", (208, 112), dict(cavlc_8x8dct=1, alpha=-2, beta=2, chroma_off=3)),
                  ("chaos_b", (96, 80), dict(chaos=True, slice_b=1, cavlc_8x8dct=1, alpha=6, beta=-4, chroma_off=-5)), ("cif_p", (352, 288), dict(qp_centre=34)))


def umh_results(o, g, pe, pr, fh, fv, fc, integ):
    """--me umh on the 160x128 pair: scattered and agreeing predictors, three ranges, full-pel and sub-pel"""
    res = []
    for me_range in (16, 24, 8):
        for spread, centre in ((48, None), (10, (-20, -12))):
            _, mis = make_me_jobs(pkg, g, seed=900 + me_range + spread, n=80, me_range=me_range, qp=(12, 26, 40), pixels=(0, 1, 2, 3, 4, 5, 6),
                                  mvp_spread=spread, centre=centre)
            for i, mi in enumerate(mis):
                mi.me_method = X.ME_UMH
                mi.b_sub8x8 = 1
                if centre is not None and i % 2:
                    for k in range(mi.i_mvc):
                        mi.mvc[k][0], mi.mvc[k][1] = mi.mvp[0] + (k % 3) - 1, mi.mvp[1] + (k % 2)
                for subme in (1, 5):
                    a = o.me_search_subpel(g, pe, [pr, fh, fv, fc], integ, mi, subme, 1)
                    res.append((me_range, subme, a.mv[0], a.mv[1], a.cost, a.cost_mv))
    return res


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    r = X.ref()
    assert r.backend == "reference"
    out = {}
    # ---- pixel metrics on seeded blocks + the max-difference patterns of checkasm.c:246-256
    rng = np.random.default_rng(2009)
    a = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    b = rng.integers(0, 256, (48, 64), dtype=np.uint8)
    yy, xx = np.mgrid[0:48, 0:64]
    pats = [(a, b), (np.zeros_like(a), np.full_like(a, 255)), ((((xx + yy) & 1) * 255).astype(np.uint8), (((xx + yy + 1) & 1) * 255).astype(np.uint8)),
            (((xx & 1) * 255).astype(np.uint8), ((yy & 1) * 255).astype(np.uint8))]
    offs = rng.integers(0, 32, (40, 4))
    vals = []
    for pa, pb in pats:
        for m in range(4):
            for ip in range(7):
                if m == X.SA8D and ip not in (0, 3):
                    continue
                for o in offs:
                    vals.append(r.pixel_cmp(m, ip, pa, 64, pb, 64, int(o[0]) * 64 + int(o[1]), int(o[2]) * 64 + int(o[3])))
    out["pixel_a"], out["pixel_b"], out["pixel_offs"], out["pixel_vals"] = a, b, offs, np.array(vals, np.int64)
    # ---- MV cost tables: sha256 of all 52 (the tables themselves are 1.7 MB)
    out["cost_mv_sha"] = np.array([sha(r.cost_mv_table(q)) for q in range(52)])
    out["cost_mv_qp26_head"] = r.cost_mv_table(26)[2 * 4 * 2048:2 * 4 * 2048 + 256]
    # ---- frame filters on a small picture (non-mod16 size on purpose) and CIF (checksums)
    for name, (w, h) in (("small", (76, 52)), ("cif", (352, 288))):
        g = r.geometry(w, h)
        pic = synth.Clip(w, h, seed=5).luma(2)
        plane = r.plane_from_picture(g, pic)
        for s8 in (0, 1):
            fh, fv, fc, integ = r.frame_filter(g, plane, s8)
            if name == "small":
                out["small_pic"] = pic
                out["small_plane"], out["small_h"], out["small_v"], out["small_c"] = plane, fh, fv, fc
                out["small_integral%d" % s8] = integ
            out["%s_filter_sha_s8%d" % (name, s8)] = np.array([sha(plane), sha(fh), sha(fv), sha(fc), sha(integ)])
        if g.mb_width % 2 == 0:
            lows = r.init_lowres(g, plane.copy())
            out["%s_lowres_sha" % name] = np.array([sha(x) for x in lows])
    # ---- motion search (ESA, TESA, DIA, HEX; sub-pel) on a 160x128 pair
    w, h = 160, 128
    g = r.geometry(w, h)
    clip = synth.Clip(w, h, seed=21)
    pe, pr = r.plane_from_picture(g, clip.luma(1)), r.plane_from_picture(g, clip.luma(0))
    fh, fv, fc, integ = r.frame_filter(g, pr, 1)
    res = []
    for method in (X.ME_DIA, X.ME_HEX, X.ME_ESA, X.ME_TESA):
        for fpel_satd in ((0, 1) if method == X.ME_TESA else (0,)):
            jobs, mis = make_me_jobs(pkg, g, seed=100 + method, n=120, me_range=16, qp=(12, 26, 40), pixels=(0, 1, 2, 3, 4, 5, 6),
                                     tesa=(method == X.ME_TESA), fpel_satd=bool(fpel_satd))
            for mi in mis:
                mi.me_method = method
                mi.b_sub8x8 = 1
                o = r.me_search_fpel(g, pe, pr, integ, mi)
                res.append((method, fpel_satd, o.mv[0], o.mv[1], o.cost, o.cost_mv))
            for subme in (2, 4, 6):
                for mi in mis[:40]:
                    mi.i_pixel = mi.i_pixel % 4
                    mi.bx, mi.by = (mi.bx // 16) * 16, (mi.by // 16) * 16
                    mi.fpel_satd = 1 if (method == X.ME_TESA) else 0
                    o = r.me_search_subpel(g, pe, [pr, fh, fv, fc], integ, mi, subme, 1)
                    res.append((method, 10 + subme, o.mv[0], o.mv[1], o.cost, o.cost_mv))
    out["me_results"] = np.array(res, np.int32)
    out["me_results_umh"] = np.array(umh_results(r, g, pe, pr, fh, fv, fc, integ), np.int32)
    # ---- lowres lookahead schedule (I, P, P dist 2, B, cached) through the reference's x264_rc_analyse_slice
    for tag, (w, h), method, satd, weighted in LOOKAHEAD_GOLDEN:
        g = r.geometry(w, h)
        planes = lowres_planes(r, g, synth.Clip(w, h, seed=31), 3)
        scal, arrs = lookahead_digest(oracle_lookahead(r, g, planes, method, 16, satd, weighted, is_ref=True), g)
        out["la_%s_scalars" % tag], out["la_%s_arrays" % tag] = scal, arrs.astype(np.int32)
    # ---- deblocking: sha256 of the three filtered planes for seeded macroblock state (x264_frame_deblock_row)
    for tag, (w, h), kw in DEBLOCK_GOLDEN:
        g = r.geometry(w, h)
        info = make_deblock_info(g, seed=100, **kw)
        y, u, v = blocky_recon(synth.Clip(w, h, seed=5), g, seed=0)
        py = r.new_plane(g)
        py.reshape(-1, g.stride)[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]] = y
        r.frame_deblock(g, info, py, u, v)
        out["deblock_%s_sha" % tag] = np.array([sha(py.reshape(-1, g.stride)[X.PADV:X.PADV + y.shape[0], X.PADH:X.PADH + y.shape[1]]), sha(u), sha(v)])
    # ---- inter residual driver (x264_macroblock_encode) and x264_macroblock_probe_skip on tiles around the skip decision
    for cqm in (0, 1):
        out["resid_cqm%d_rows" % cqm], out["skip_cqm%d" % cqm] = residual_digest(r, skip_probe_cases(80 + cqm, 300), cqm)
    # ---- Intra16x16 + chroma candidate costs through the reference's predict_16x16 / predict_8x8c tables and mbcmp functions
    out["intra_costs"] = intra_digest(r, intra_cases(91, 400))
    out["lambda2_tab"] = np.array([r.lib.xo_lambda2(q) for q in range(52)], np.int32)
    np.savez_compressed(os.path.join(HERE, "reference_vectors.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_vectors.npz"), os.path.getsize(os.path.join(HERE, "reference_vectors.npz")), "bytes")


if __name__ == "__main__":
    main()
