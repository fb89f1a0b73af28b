"""gpu: batched ESA motion search through the C ABI vs the oracle (same seeded inputs), bit-exact"""
import numpy as np
import pytest
import xo_api as X
from helpers import make_me_jobs, oracle_me, make_mb_jobs, mb_jobs_to_block_jobs, block_jobs_to_mis

pytestmark = pytest.mark.gpu


def _setup(pkg, ctx, port, w, h, seed, sub8=0):
    from x264_vs2008_b200 import synth
    clip = synth.Clip(w, h, seed=seed)
    g = port.geometry(w, h)
    fenc, fref = ctx.frame(w, h, 0), ctx.frame(w, h, 0)
    fenc.upload(clip.luma(1)); fenc.expand_border()
    fref.upload(clip.luma(0)); fref.expand_border()
    pe, pr = port.plane_from_picture(g, clip.luma(1)), port.plane_from_picture(g, clip.luma(0))
    return g, fenc, fref, pe, pr


@pytest.mark.parametrize("w,h,me_range,pixels", [
    (128, 96, 16, (0, 1, 2, 3)),
    (352, 288, 16, (0, 1, 2, 3)),
    (352, 288, 8, (0, 3)),
    (352, 288, 24, (0, 1, 2, 3, 4, 5, 6)),
    (352, 288, 40, (0, 3, 6)),       # window wider than one warp: column chunks
    (100, 70, 16, (0, 1, 2, 3, 4, 5, 6)),  # non-mod16 picture
])
def test_esa_matches_oracle(pkg, ctx, port, w, h, me_range, pixels):
    g, fenc, fref, pe, pr = _setup(pkg, ctx, port, w, h, seed=w + me_range)
    jobs, mis = make_me_jobs(pkg, g, seed=me_range, n=600, me_range=me_range, qp=(12, 20, 26, 32, 45), pixels=pixels)
    res = ctx.me_search(fenc, fref, me_range, jobs)
    want = oracle_me(port, g, pe, pr, None, mis)
    got = [(int(r["bmx"]), int(r["bmy"]), int(r["bcost"])) for r in res]
    bad = [i for i in range(len(got)) if got[i] != want[i]]
    assert not bad, (len(bad), bad[:5], [got[i] for i in bad[:5]], [want[i] for i in bad[:5]])
    # the seed (predictor stage) must match too
    for r, mi in list(zip(res, mis))[:200]:
        o = port.me_search_fpel(g, pe, pr, None, mi)
        assert (int(r["seed_mx"]), int(r["seed_my"]), int(r["seed_cost"])) == (o.seed_mx, o.seed_my, o.seed_cost)
    fenc.close(); fref.close()


def test_esa_flat_content_ties(pkg, ctx, port):
    """all-equal SADs: the winner is decided purely by the raster-order / strict-'<' rule and the MV cost"""
    w, h = 160, 128
    g = port.geometry(w, h)
    for val_e, val_r in ((0, 0), (255, 0), (17, 17)):
        fenc, fref = ctx.frame(w, h, 0), ctx.frame(w, h, 0)
        fenc.upload(np.full((h, w), val_e, np.uint8)); fenc.expand_border()
        fref.upload(np.full((h, w), val_r, np.uint8)); fref.expand_border()
        pe = port.plane_from_picture(g, np.full((h, w), val_e, np.uint8))
        pr = port.plane_from_picture(g, np.full((h, w), val_r, np.uint8))
        jobs, mis = make_me_jobs(pkg, g, seed=val_e, n=200, me_range=16, qp=(0, 26, 51))
        res = ctx.me_search(fenc, fref, 16, jobs)
        want = oracle_me(port, g, pe, pr, None, mis)
        assert [(int(r["bmx"]), int(r["bmy"]), int(r["bcost"])) for r in res] == want
        fenc.close(); fref.close()


def test_esa_seeded_jobs(pkg, ctx, port):
    """X264_CUDA_ME_SEEDED: caller supplies bmx,bmy,bcost (the sub-pel predictor flow of me.c:189-205)"""
    w, h = 352, 288
    g, fenc, fref, pe, pr = _setup(pkg, ctx, port, w, h, seed=77)
    jobs, mis = make_me_jobs(pkg, g, seed=1, n=300, me_range=16, qp=26)
    base = ctx.me_search(fenc, fref, 16, jobs)
    j2 = jobs.copy()
    j2["flags"] |= pkg.ME_SEEDED
    j2["seed_mv"][:, 0], j2["seed_mv"][:, 1], j2["seed_cost"] = base["seed_mx"], base["seed_my"], base["seed_cost"]
    j2["i_mvc"] = 0
    again = ctx.me_search(fenc, fref, 16, j2)
    assert np.array_equal(again["bmx"], base["bmx"]) and np.array_equal(again["bmy"], base["bmy"])
    assert np.array_equal(again["bcost"], base["bcost"])
    fenc.close(); fref.close()


def test_esa_1080p_full_frame_property(pkg, ctx, port):
    """BASELINE config 2 at full size: every MB of a 1080p pair, 16x16 ESA.  Properties: (i) shifting the
    reference by a known whole-pel vector makes every interior MB find exactly that vector with SAD 0 when the
    predictor points nowhere near it; (ii) a spot-check of 400 random jobs against the oracle."""
    w, h = 1920, 1080
    rng = np.random.default_rng(4)
    big = rng.integers(0, 256, (h + 32, w + 32), dtype=np.uint8)
    dx, dy = 7, -5
    cur = np.ascontiguousarray(big[16:16 + h, 16:16 + w])
    refp = np.ascontiguousarray(big[16 - dy:16 - dy + h, 16 - dx:16 - dx + w])  # ref(x,y) = cur(x-dx, y-dy)
    g = port.geometry(w, h)
    fenc, fref = ctx.frame(w, h, 0), ctx.frame(w, h, 0)
    fenc.upload(cur); fenc.expand_border()
    fref.upload(refp); fref.expand_border()
    n = g.mb_width * g.mb_height
    jobs = np.zeros(n, pkg.ME_JOB)
    for mby in range(g.mb_height):
        for mbx in range(g.mb_width):
            j = jobs[mby * g.mb_width + mbx]
            mnf, mxf, _, _ = X.mv_limits_fpel(g, mbx, mby)
            j["bx"], j["by"], j["qp"] = mbx * 16, mby * 16, 26
            j["mv_min_fpel"], j["mv_max_fpel"] = mnf, mxf
    res = ctx.me_search(fenc, fref, 16, jobs)
    res2d = res.reshape(g.mb_height, g.mb_width)
    inner = res2d[2:-2, 2:-2]
    assert np.all(inner["bmx"] == dx) and np.all(inner["bmy"] == dy)
    lam_bits = pkg.host_cost_mv(26)[2 * 4 * 2048 + 4 * dx] + pkg.host_cost_mv(26)[2 * 4 * 2048 + 4 * abs(dy)]
    assert np.all(inner["bcost"] == lam_bits)
    pe, pr = port.plane_from_picture(g, cur), port.plane_from_picture(g, refp)
    pick = rng.choice(n, 400, replace=False)
    for i in pick:
        mi = X.MeIn()
        mi.me_method, mi.me_range, mi.qp, mi.i_pixel = X.ME_ESA, 16, 26, 0
        mi.bx, mi.by = int(jobs[i]["bx"]), int(jobs[i]["by"])
        for k in range(2):
            mi.mv_min_fpel[k], mi.mv_max_fpel[k] = int(jobs[i]["mv_min_fpel"][k]), int(jobs[i]["mv_max_fpel"][k])
        o = port.me_search_fpel(g, pe, pr, None, mi)
        assert (int(res[i]["bmx"]), int(res[i]["bmy"]), int(res[i]["bcost"])) == (o.bmx, o.bmy, o.bcost)
    fenc.close(); fref.close()


@pytest.mark.parametrize("w,h,me_range", [(352, 288, 16), (352, 288, 8), (352, 288, 24), (128, 96, 16), (100, 70, 16), (640, 360, 40)])
def test_mb_search_matches_oracle_and_block_search(pkg, ctx, port, w, h, me_range):
    """macroblock-batched search (nine partitions, union window) == nine independent x264_me_search_ref"""
    g, fenc, fref, pe, pr = _setup(pkg, ctx, port, w, h, seed=3 * w + me_range)
    mbjobs = make_mb_jobs(pkg, g, seed=me_range, n=250, qp=(12, 26, 38))
    res = ctx.me_search_mb(fenc, fref, me_range, mbjobs)
    bjobs, idx = mb_jobs_to_block_jobs(pkg, mbjobs)
    outs = port.me_search_fpel_batch(g, pe, pr, None, block_jobs_to_mis(bjobs, me_range))
    bres = ctx.me_search(fenc, fref, me_range, bjobs)
    bad = []
    for k, (i, p) in enumerate(idx):
        r = res[i]["part"][p]
        got = (int(r["bmx"]), int(r["bmy"]), int(r["bcost"]), int(r["seed_mx"]), int(r["seed_my"]), int(r["seed_cost"]))
        o = outs[k]
        want = (o.bmx, o.bmy, o.bcost, o.seed_mx, o.seed_my, o.seed_cost)
        blk = bres[k]
        if got != want or got[:3] != (int(blk["bmx"]), int(blk["bmy"]), int(blk["bcost"])):
            bad.append((i, p, got, want))
    assert not bad, (len(bad), bad[:5])
    # masked-out partitions are flagged
    for i, mj in enumerate(mbjobs):
        for p in range(9):
            if not (int(mj["part_mask"]) >> p) & 1:
                assert int(res[i]["part"][p]["bcost"]) == -1
    fenc.close(); fref.close()


def test_blocking_wait_mode_gives_the_same_results(pkg, ctx, port):
    """x264_cuda_set_blocking_wait: the host thread sleeps on a blocking-sync event instead of spinning; results are those of the default mode"""
    g, fenc, fref, pe, pr = _setup(pkg, ctx, port, 352, 288, seed=77)
    mbjobs = make_mb_jobs(pkg, g, seed=5, n=120, qp=(20, 30))
    want = ctx.me_search_mb(fenc, fref, 16, mbjobs)
    ctx.set_blocking_wait(1)
    try:
        got = ctx.me_search_mb(fenc, fref, 16, mbjobs)
        bjobs, _ = mb_jobs_to_block_jobs(pkg, mbjobs)
        blk = ctx.me_search(fenc, fref, 16, bjobs)
    finally:
        ctx.set_blocking_wait(0)
    assert got.tobytes() == want.tobytes()
    assert np.array_equal(blk, ctx.me_search(fenc, fref, 16, bjobs))
    fenc.close(); fref.close()
