"""not-gpu: the host-side ESA replay (x264_cuda_host_esa_replay, product code in host/x264_cuda_host.c) on grids computed with numpy
must reproduce the oracle's x264_me_search_ref full-pel result (and, where oracle/_ref exists, the reference's)."""
import numpy as np
import xo_api as X
from helpers import make_me_jobs, oracle_me

PARTS = [(0, 0, 0), (1, 0, 0), (1, 0, 8), (2, 0, 0), (2, 8, 0), (3, 0, 0), (3, 8, 0), (3, 0, 8), (3, 8, 8)]


def numpy_grid(g, pe, pr, mbx, mby, ip, ox, oy, cx, cy, radius, mn, mx, gw, gh):
    E = pe.reshape(-1, g.stride).astype(np.int32)
    R = pr.reshape(-1, g.stride).astype(np.int32)
    bw, bh = X.BLK_W[ip], X.BLK_H[ip]
    y0, x0 = X.PADV + 16 * mby + oy, X.PADH + 16 * mbx + ox
    blk = E[y0:y0 + bh, x0:x0 + bw]
    out = np.full((gh, gw), 0xffff, np.uint16)
    for j in range(gh):
        for i in range(gw):
            vx, vy = cx - radius + i, cy - radius + j
            if mn[0] <= vx <= mx[0] + 3 and mn[1] <= vy <= mx[1]:
                out[j, i] = np.abs(blk - R[y0 + vy:y0 + vy + bh, x0 + vx:x0 + vx + bw]).sum()
    return out


def test_esa_replay_on_numpy_grids(pkg, port):
    from x264_vs2008_b200 import synth
    w, h = 160, 128
    me_range, radius = 8, 12
    g = port.geometry(w, h)
    clip = synth.Clip(w, h, seed=21)
    pe, pr = port.plane_from_picture(g, clip.luma(1)), port.plane_from_picture(g, clip.luma(0))
    gw, gh = pkg.grid_w(radius), pkg.grid_h(radius)
    integ = X.ref().frame_filter(g, pr, 0)[3] if X.have_ref() else None  # the reference's ESA prunes with ADS on the integral image
    rng = np.random.default_rng(5)
    n_ok = n_out = 0
    for t in range(60):
        mbx, mby = int(rng.integers(0, g.mb_width)), int(rng.integers(0, g.mb_height))
        p = int(rng.integers(0, 9))
        ip, ox, oy = PARTS[p]
        qp = int(rng.choice([20, 26, 32]))
        mn, mx, _, _ = X.mv_limits_fpel(g, mbx, mby)
        cx, cy = int(np.clip(-5, mn[0], mx[0])), int(np.clip(-3, mn[1], mx[1]))
        jobs, mis = make_me_jobs(pkg, g, seed=t, n=1, me_range=me_range, qp=qp, pixels=(ip,), mvp_spread=16)
        job, mi = jobs[0], mis[0]
        job["bx"], job["by"] = 16 * mbx + ox, 16 * mby + oy
        mi.bx, mi.by = int(job["bx"]), int(job["by"])
        job["mv_min_fpel"], job["mv_max_fpel"] = mn, mx
        for k in range(2):
            mi.mv_min_fpel[k], mi.mv_max_fpel[k] = int(mn[k]), int(mx[k])
        job["mvp"][0], job["mvp"][1] = job["mvp"][0] - 20, job["mvp"][1] - 12
        mi.mvp[0], mi.mvp[1] = int(job["mvp"][0]), int(job["mvp"][1])
        for k in range(int(job["i_mvc"])):
            v = [0, 0] if rng.integers(0, 6) == 0 else [int(rng.integers(-12, 13)) - 20, int(rng.integers(-12, 13)) - 12]
            job["mvc"][k] = v
            mi.mvc[k][0], mi.mvc[k][1] = v
        grid = numpy_grid(g, pe, pr, mbx, mby, ip, ox, oy, cx, cy, radius, mn, mx, gw, gh)
        res = pkg.host_esa_replay(grid, radius, cx, cy, job, me_range, pkg.host_cost_mv(qp))
        if res is None:
            n_out += 1
            continue
        want = oracle_me(port, g, pe, pr, None, [mi])[0]
        assert (int(res["bmx"]), int(res["bmy"]), int(res["bcost"])) == tuple(want), (t, p, res, want)
        if X.have_ref():
            r = X.ref().me_search_fpel(g, pe, pr, integ, mi)
            mv = np.zeros(2, np.int16); cost = np.zeros(1, np.int32); cmv = np.zeros(1, np.int32)
            pkg.lib().x264_cuda_me_finish(np.ascontiguousarray(job).ctypes.data, np.ascontiguousarray(res).ctypes.data, pkg.host_cost_mv(qp).ctypes.data,
                                          int(mi.mv_max_spel[1]), mv.ctypes.data, cost.ctypes.data, cmv.ctypes.data)
            assert (int(mv[0]), int(mv[1]), int(cost[0]), int(cmv[0])) == (r.mv[0], r.mv[1], r.cost, r.cost_mv), (t, p)
        n_ok += 1
    assert n_ok >= 40, (n_ok, n_out)
