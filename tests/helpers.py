"""shared helpers for tests/, smoke() and bench.py's checking legs (oracle side only)"""
import numpy as np
import xo_api as X


def pad_view(g, flat, lowres=False):
    """oracle plane buffer (flat, reference stride) -> 2-D padded view (rows -32.., cols -32..) cropped to the
    width the device download uses (w16 + 64)"""
    stride = g.stride_lowres if lowres else g.stride
    lines = g.lines_lowres if lowres else g.lines
    w = (g.width_lowres if lowres else g.mb_width * 16) + 2 * X.PADH
    return flat.reshape(lines + 2 * X.PADV, stride)[:, :w]


def make_me_jobs(pkg, g, seed, n, me_range, qp, pixels=(0, 1, 2, 3), mvp_spread=48, tesa=False, fpel_satd=False,
                 centre=None):
    """n seeded search jobs over random macroblocks: returns (numpy ME_JOB array, list of xo MeIn)"""
    rng = np.random.default_rng(seed)
    jobs = np.zeros(n, pkg.ME_JOB)
    mis = []
    for i in range(n):
        ip = int(pixels[int(rng.integers(0, len(pixels)))])
        mbx, mby = int(rng.integers(0, g.mb_width)), int(rng.integers(0, g.mb_height))
        bw, bh = X.BLK_W[ip], X.BLK_H[ip]
        bx = mbx * 16 + int(rng.integers(0, 16 // bw)) * bw
        by = mby * 16 + int(rng.integers(0, 16 // bh)) * bh
        mnf, mxf, mns, mxs = X.mv_limits_fpel(g, mbx, mby)
        cx, cy = centre if centre is not None else (0, 0)
        mvp = [int(rng.integers(-mvp_spread, mvp_spread + 1)) + cx, int(rng.integers(-mvp_spread, mvp_spread + 1)) + cy]
        nmvc = int(rng.integers(0, 6))
        mi = X.MeIn()
        mi.me_method = X.ME_TESA if tesa else X.ME_ESA
        mi.me_range = me_range
        mi.qp = qp if isinstance(qp, int) else int(qp[int(rng.integers(0, len(qp)))])
        mi.fpel_satd = int(fpel_satd)
        mi.i_pixel = ip
        mi.bx, mi.by = bx, by
        for k in range(2):
            mi.mv_min_fpel[k], mi.mv_max_fpel[k] = mnf[k], mxf[k]
            mi.mv_min_spel[k], mi.mv_max_spel[k] = mns[k], mxs[k]
            mi.mvp[k] = mvp[k]
        mi.i_mvc = nmvc
        j = jobs[i]
        j["bx"], j["by"], j["i_pixel"], j["qp"], j["i_mvc"] = bx, by, ip, mi.qp, nmvc
        j["flags"] = (pkg.ME_TESA if tesa else 0) | (pkg.ME_FPEL_SATD if fpel_satd else 0)
        j["mvp"] = mvp
        j["mv_min_fpel"], j["mv_max_fpel"] = mnf, mxf
        for k in range(nmvc):
            v = [int(rng.integers(-120, 121)) + cx, int(rng.integers(-120, 121)) + cy]
            if rng.integers(0, 8) == 0:
                v = [0, 0]
            mi.mvc[k][0], mi.mvc[k][1] = v
            j["mvc"][k] = v
        mis.append(mi)
    return jobs, mis


def oracle_me(o, g, fenc, fref, integral, mis):
    """[(bmx,bmy,bcost)] from the oracle; falls back to (mv>>2, cost-adjusted) when the backend hides fpel state"""
    out = []
    for mi in mis:
        r = o.me_search_fpel(g, fenc, fref, integral, mi)
        out.append((r.bmx, r.bmy, r.bcost))
    return out
