"""gpu: candidate grids (x264_cuda_sad_grid) and the host replay of x264_me_search_ref's ESA branch on them:
every grid value == x264_pixel_sad_* of the oracle; replayed (mv, cost) == x264_me_search_ref for seeded (mvp, mvc, qp in {20,26,32})."""
import numpy as np
import pytest
import xo_api as X
from helpers import make_me_jobs, oracle_me

pytestmark = pytest.mark.gpu

PARTS = [(0, 0, 0), (1, 0, 0), (1, 0, 8), (2, 0, 0), (2, 8, 0), (3, 0, 0), (3, 8, 0), (3, 0, 8), (3, 8, 8)]  # (i_pixel, x offset, y offset)


def _frames(pkg, ctx, port, w, h, seed):
    from x264_vs2008_b200 import synth
    clip = synth.Clip(w, h, seed=seed)
    g = port.geometry(w, h)
    fenc, fref = ctx.frame(w, h, 0), ctx.frame(w, h, 0)
    fenc.upload(clip.luma(1)); fenc.expand_border()
    fref.upload(clip.luma(0)); fref.expand_border()
    pe, pr = port.plane_from_picture(g, clip.luma(1)), port.plane_from_picture(g, clip.luma(0))
    return g, fenc, fref, pe, pr


@pytest.mark.parametrize("radius", [16, 20, 7])
def test_grid_values(pkg, ctx, port, radius):
    w, h = 208, 144
    g, fenc, fref, pe, pr = _frames(pkg, ctx, port, w, h, 11)
    rng = np.random.default_rng(radius)
    n = g.mb_width * g.mb_height
    jobs = np.zeros(n, pkg.GRID_JOB)
    for i in range(n):
        mbx, mby = i % g.mb_width, i // g.mb_width
        mn, mx, _, _ = X.mv_limits_fpel(g, mbx, mby)
        jobs[i]["mb_x"], jobs[i]["mb_y"] = mbx, mby
        jobs[i]["cx"], jobs[i]["cy"] = int(np.clip(rng.integers(-12, 13), mn[0], mx[0])), int(np.clip(rng.integers(-12, 13), mn[1], mx[1]))
        jobs[i]["mv_min_fpel"], jobs[i]["mv_max_fpel"] = mn, mx
        jobs[i]["part_mask"] = 511 if i % 7 else int(rng.integers(1, 512))
    grid = ctx.sad_grid(fenc, fref, radius, jobs)
    E = pe.reshape(-1, g.stride).astype(np.int32)
    R = pr.reshape(-1, g.stride).astype(np.int32)
    GW, GH = pkg.grid_w(radius), pkg.grid_h(radius)
    bw, bh = [16, 16, 8, 8], [16, 8, 16, 8]
    n_checked = 0
    for i in rng.choice(n, 24, replace=False):
        j = jobs[i]
        for p, (ip, ox, oy) in enumerate(PARTS):
            y0, x0 = X.PADV + 16 * int(j["mb_y"]) + oy, X.PADH + 16 * int(j["mb_x"]) + ox
            blk = E[y0:y0 + bh[ip], x0:x0 + bw[ip]]
            for jj in range(GH):
                my = int(j["cy"]) - radius + jj
                for ii in range(0, GW, 3 if jj % 4 else 1):
                    mx = int(j["cx"]) - radius + ii
                    ok = (int(j["part_mask"]) >> p & 1) and j["mv_min_fpel"][0] <= mx <= j["mv_max_fpel"][0] + 3 and j["mv_min_fpel"][1] <= my <= j["mv_max_fpel"][1]
                    want = 0xffff
                    if ok:
                        want = int(np.abs(blk - R[y0 + my:y0 + my + bh[ip], x0 + mx:x0 + mx + bw[ip]]).sum())
                    assert int(grid[i, p, jj, ii]) == want, (i, p, jj, ii, mx, my)
                    n_checked += 1
    # the same numbers through the oracle's x264_pixel_sad_* on a few positions
    for i in (0, n // 2, n - 1):
        j = jobs[i]
        for p, (ip, ox, oy) in enumerate(PARTS):
            if not (int(j["part_mask"]) >> p & 1):
                continue
            ii, jj = radius, radius  # the centre itself
            off_e = g.origin + (16 * int(j["mb_y"]) + oy) * g.stride + 16 * int(j["mb_x"]) + ox
            off_r = off_e + int(j["cy"]) * g.stride + int(j["cx"])
            assert int(grid[i, p, jj, ii]) == port.pixel_cmp(X.SAD, ip, pe, g.stride, pr, g.stride, off_e, off_r)
    assert n_checked > 20000
    fenc.close(); fref.close()


@pytest.mark.parametrize("qp", [20, 26, 32])
def test_esa_replay(pkg, ctx, port, qp):
    """sequential-predictor use: the grid is computed around a guessed centre, the exact mvp/mvc arrive later on the host"""
    w, h = 320, 192
    me_range, radius = 16, 22
    g, fenc, fref, pe, pr = _frames(pkg, ctx, port, w, h, 23)
    n = g.mb_width * g.mb_height
    gj = np.zeros(n, pkg.GRID_JOB)
    rng = np.random.default_rng(qp)
    for i in range(n):
        mbx, mby = i % g.mb_width, i // g.mb_width
        mn, mx, _, _ = X.mv_limits_fpel(g, mbx, mby)
        gj[i]["mb_x"], gj[i]["mb_y"], gj[i]["part_mask"] = mbx, mby, 511
        gj[i]["cx"], gj[i]["cy"] = int(np.clip(-5, mn[0], mx[0])), int(np.clip(-3, mn[1], mx[1]))  # the clip's global motion as the guess
        gj[i]["mv_min_fpel"], gj[i]["mv_max_fpel"] = mn, mx
    grid = ctx.sad_grid(fenc, fref, radius, gj)
    table = pkg.host_cost_mv(qp)
    n_ok = n_out = 0
    for i in rng.choice(n, 80, replace=False):
        mbx, mby = int(gj[i]["mb_x"]), int(gj[i]["mb_y"])
        for p, (ip, ox, oy) in enumerate(PARTS):
            jobs, mis = make_me_jobs(pkg, g, seed=1000 * i + p, n=1, me_range=me_range, qp=qp, pixels=(ip,), mvp_spread=24)
            job, mi = jobs[0], mis[0]
            job["bx"], job["by"] = 16 * mbx + ox, 16 * mby + oy
            mi.bx, mi.by = int(job["bx"]), int(job["by"])
            job["mv_min_fpel"], job["mv_max_fpel"] = gj[i]["mv_min_fpel"], gj[i]["mv_max_fpel"]
            for k in range(2):
                mi.mv_min_fpel[k], mi.mv_max_fpel[k] = int(gj[i]["mv_min_fpel"][k]), int(gj[i]["mv_max_fpel"][k])
            job["mvp"][0], job["mvp"][1] = job["mvp"][0] - 20, job["mvp"][1] - 12  # predictors scattered around the true motion
            mi.mvp[0], mi.mvp[1] = int(job["mvp"][0]), int(job["mvp"][1])
            for k in range(int(job["i_mvc"])):  # neighbours' vectors: near the motion too (one of them may be zero)
                v = [0, 0] if rng.integers(0, 6) == 0 else [int(rng.integers(-28, 29)) - 20, int(rng.integers(-28, 29)) - 12]
                job["mvc"][k] = v
                mi.mvc[k][0], mi.mvc[k][1] = v
            res = pkg.host_esa_replay(grid[i, p], radius, int(gj[i]["cx"]), int(gj[i]["cy"]), job, me_range, table)
            if res is None:
                n_out += 1
                continue
            want = oracle_me(port, g, pe, pr, None, [mi])[0]
            assert (int(res["bmx"]), int(res["bmy"]), int(res["bcost"])) == tuple(want), (i, p, res, want)
            n_ok += 1
    assert n_ok > 8 * n_out and n_ok > 400, (n_ok, n_out)
    fenc.close(); fref.close()


@pytest.mark.parametrize("size,radius", [((208, 144), 24), ((1920, 1080), 24), ((352, 288), 40)])
def test_quad_grid_equals_partition_grid(pkg, ctx, port, size, radius):
    """the quadrant grids the live encoder reads (x264_cuda_sad_grid_quad: TL, TR, BL, BR per position, interleaved) against the
    nine-plane grids checked above: planes 5..8 are the quadrants themselves and every other partition is a sum of them"""
    w, h = size
    g, fenc, fref, pe, pr = _frames(pkg, ctx, port, w, h, 13)
    rng = np.random.default_rng(radius)
    n_all = g.mb_width * g.mb_height
    pick = rng.choice(n_all, min(n_all, 160), replace=False)
    jobs = np.zeros(len(pick), pkg.GRID_JOB)
    for k, i in enumerate(pick):
        mbx, mby = int(i % g.mb_width), int(i // g.mb_width)
        mn, mx, _, _ = X.mv_limits_fpel(g, mbx, mby)
        jobs[k]["mb_x"], jobs[k]["mb_y"] = mbx, mby
        jobs[k]["cx"], jobs[k]["cy"] = int(np.clip(rng.integers(-20, 21), mn[0], mx[0])), int(np.clip(rng.integers(-20, 21), mn[1], mx[1]))
        jobs[k]["mv_min_fpel"], jobs[k]["mv_max_fpel"] = mn, mx
        jobs[k]["part_mask"] = 511
    nine = ctx.sad_grid(fenc, fref, radius, jobs).astype(np.int64)
    quad = ctx.sad_grid_quad(fenc, fref, radius, jobs).astype(np.int64)
    invalid = nine[:, 5] == 0xffff
    assert np.array_equal(invalid, quad[..., 0] == 0xffff) and invalid.any() and not invalid.all()
    for q in range(4):
        assert np.array_equal(quad[..., q], nine[:, 5 + q])
    ok = ~invalid
    sums = {0: quad.sum(-1), 1: quad[..., 0] + quad[..., 1], 2: quad[..., 2] + quad[..., 3], 3: quad[..., 0] + quad[..., 2], 4: quad[..., 1] + quad[..., 3]}
    for p, v in sums.items():
        assert np.array_equal(v[ok], nine[:, p][ok]), p
    fenc.close(); fref.close()
