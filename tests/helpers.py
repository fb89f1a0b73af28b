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


def make_mb_jobs(pkg, g, seed, n, qp, mvp_spread=40, mvc_spread=60, centre=(0, 0), mask_random=True):
    """n seeded macroblock jobs (nine partition searches each) -> (ME_MB_JOB array)"""
    rng = np.random.default_rng(seed)
    jobs = np.zeros(n, pkg.ME_MB_JOB)
    for i in range(n):
        j = jobs[i]
        mbx, mby = int(rng.integers(0, g.mb_width)), int(rng.integers(0, g.mb_height))
        mnf, mxf, _, _ = X.mv_limits_fpel(g, mbx, mby)
        j["mb_x"], j["mb_y"] = mbx, mby
        j["qp"] = qp if isinstance(qp, int) else int(qp[int(rng.integers(0, len(qp)))])
        j["part_mask"] = int(rng.integers(1, 512)) if (mask_random and i % 4 == 0) else 511
        j["mv_min_fpel"], j["mv_max_fpel"] = mnf, mxf
        base = [int(rng.integers(-mvp_spread, mvp_spread + 1)) + centre[0], int(rng.integers(-mvp_spread, mvp_spread + 1)) + centre[1]]
        for p in range(pkg.ME_MB_PARTS):
            # partitions of one MB have correlated predictors (as in the encoder) with occasional outliers
            jit = 6 if rng.integers(0, 6) else 70
            j["mvp"][p] = [base[0] + int(rng.integers(-jit, jit + 1)), base[1] + int(rng.integers(-jit, jit + 1))]
            # every third job: the 16x16 search carries 5..11 predictors like x264_mb_predict_mv_ref16x16's list; the ones beyond four live
            # in the fourth slot of partitions 1..7 (X264_CUDA_ME_MB_MVC16), which then use at most three themselves (the last 8x8 may use four)
            wide = i % 3 == 1
            nm = int(rng.integers(5, 12)) if (wide and p == 0) else int(rng.integers(0, pkg.ME_MB_MVC + (0 if (wide and p < 8) else 1)))
            j["i_mvc"][p] = nm
            for k in range(nm):
                v = [base[0] + int(rng.integers(-mvc_spread, mvc_spread + 1)), base[1] + int(rng.integers(-mvc_spread, mvc_spread + 1))]
                if rng.integers(0, 8) == 0:
                    v = [0, 0]
                if k < pkg.ME_MB_MVC:
                    j["mvc"][p][k] = v
                else:
                    j["mvc"][k - pkg.ME_MB_MVC + 1][pkg.ME_MB_MVC - 1] = v
    return jobs


def mb_jobs_to_block_jobs(pkg, mbjobs):
    """expand macroblock jobs into the equivalent per-block ME_JOB list (masked partitions skipped);
    returns (ME_JOB array, list of (mb index, partition))"""
    out, idx = [], []
    for i, mj in enumerate(mbjobs):
        for p, (ip, ox, oy) in enumerate(pkg.ME_MB_PART_GEOM):
            if not (int(mj["part_mask"]) >> p) & 1:
                continue
            j = np.zeros((), pkg.ME_JOB)
            j["bx"], j["by"], j["i_pixel"], j["qp"] = int(mj["mb_x"]) * 16 + ox, int(mj["mb_y"]) * 16 + oy, ip, mj["qp"]
            j["i_mvc"] = mj["i_mvc"][p]
            j["mvp"] = mj["mvp"][p]
            j["mv_min_fpel"], j["mv_max_fpel"] = mj["mv_min_fpel"], mj["mv_max_fpel"]
            j["mvc"][:pkg.ME_MB_MVC] = mj["mvc"][p]
            if p == 0:
                for k in range(pkg.ME_MB_MVC, int(mj["i_mvc"][0])):
                    j["mvc"][k] = mj["mvc"][k - pkg.ME_MB_MVC + 1][pkg.ME_MB_MVC - 1]
            out.append(j)
            idx.append((i, p))
    return np.array(out, pkg.ME_JOB), idx


def block_jobs_to_mis(jobs, me_range, method=X.ME_ESA):
    arr = (X.MeIn * len(jobs))()
    for i, j in enumerate(jobs):
        m = arr[i]
        m.me_method, m.me_range, m.qp, m.i_pixel = method, me_range, int(j["qp"]), int(j["i_pixel"])
        m.bx, m.by, m.i_mvc = int(j["bx"]), int(j["by"]), int(j["i_mvc"])
        for k in range(2):
            m.mv_min_fpel[k], m.mv_max_fpel[k], m.mvp[k] = int(j["mv_min_fpel"][k]), int(j["mv_max_fpel"][k]), int(j["mvp"][k])
            m.mv_max_spel[k] = 1 << 20
        for c in range(m.i_mvc):
            m.mvc[c][0], m.mvc[c][1] = int(j["mvc"][c][0]), int(j["mvc"][c][1])
    return arr


# ---------------- lowres lookahead (S/encoder/slicetype.c:43-355) ----------------
# The evaluation schedule the slicetype decision produces for a 3-frame window: I cost of frame 0 and 1, P(0->1), P(0->2),
# B(0,1,2) with searches, then the same B again with cached vectors (do_search 0) and P(0->1) with cached intra costs.
LOOKAHEAD_SCHEDULE = (
    # (name, fenc, p0, p1, b, do_search, b_intra_calculated)
    ("I0", 0, 0, 0, 0, (0, 0), 0),
    ("P01", 1, 0, 1, 1, (1, 0), 0),
    ("P02", 2, 0, 2, 2, (1, 0), 0),
    ("B012", 1, 0, 2, 1, (0, 1), 1),   # list0 vectors of frame 1 at distance 1 are cached from P01
    ("B012_cached", 1, 0, 2, 1, (0, 0), 1),
    ("P01_cached", 1, 0, 1, 1, (0, 0), 1),
)


def lowres_planes(o, g, clip, n_frames):
    """[frame][4] padded lowres planes built by oracle `o` from synthetic luma"""
    out = []
    for i in range(n_frames):
        p = o.plane_from_picture(g, clip.luma(i))
        out.append(o.init_lowres(g, p))
    return out


def oracle_lookahead(o, g, planes, me_method=X.ME_HEX, me_range=16, mbcmp_satd=1, weighted=0, is_ref=False, vbv=False, inv_qscale=None):
    """runs LOOKAHEAD_SCHEDULE on oracle `o`; returns list of (name, score, intra_mbs, intra_cost_sum, mvs0, costs0, mvs1, costs1, intra,
    row_satd, score_aq).  Frame state: per frame, per (list, dist) arrays, as x264_frame_t keeps them (S/common/frame.h:65-74).
    vbv: the rc.i_vbv_buffer_size form (slicetype.c:300-316); inv_qscale: per-frame list of uint16[n_mb] (AQ on) or None."""
    n = g.mb_width * g.mb_height
    st = [{"mvs": np.zeros((2, 3, n, 2), np.int16), "costs": np.zeros((2, 3, n), np.int32), "intra": np.zeros(n, np.uint16)} for _ in planes]
    res = []
    for name, fe, p0, p1, b, ds, bic in LOOKAHEAD_SCHEDULE:
        d0, d1 = max(b - p0 - 1, 0), max(p1 - b - 1, 0)
        s = st[fe]
        state = {"mvs0": s["mvs"][0, d0], "costs0": s["costs"][0, d0], "mvs1": s["mvs"][1, d1], "costs1": s["costs"][1, d1],
                 "intra": s["intra"], "ref1_mvs": st[p1]["mvs"][0, max(p1 - p0 - 1, 0)].copy()}
        row_satd = np.zeros(g.mb_height, np.int32)
        out = o.lowres_frame_cost(g, planes[b], planes[p0], planes[p1], p0, p1, b, state, me_method=me_method, me_range=me_range,
                                  mbcmp_satd=mbcmp_satd, weighted=weighted, do_search=ds, b_intra_calculated=bic, vbv=vbv,
                                  inv_qscale=inv_qscale[fe] if inv_qscale is not None else None, row_satd=row_satd)
        score = out.score
        if not is_ref and b < p1:
            score = score * 100 // 120   # the reference stores B scores scaled (slicetype.c:338-339, i_bframe_bias 0)
        res.append((name, score, out.intra_mbs if b == p1 else 0, out.intra_cost_sum if (b == p1 and p0 != p1) else 0,  # for I the reference overwrites i_cost_est[0][0] with the score
                    state["mvs0"].copy(), state["costs0"].copy(), state["mvs1"].copy(), state["costs1"].copy(), s["intra"].copy(),
                    row_satd, out.score_aq))
    return res


def lookahead_digest(res, g, all_blocks=False):
    """compact, comparable form: scores + interior arrays (all_blocks: every block — the VBV form evaluates the frame edge too)"""
    W, H = g.mb_width, g.mb_height
    small = W <= 2 or H <= 2 or all_blocks
    m = np.zeros((H, W), bool)
    if small:
        m[:] = True
    else:
        m[1:H - 1, 1:W - 1] = True
    m = m.ravel()
    scal = np.array([(r[1], r[2], r[3]) for r in res], np.int64)
    arrs = [np.concatenate([r[4][m].ravel().astype(np.int64), r[5][m], r[6][m].ravel().astype(np.int64), r[7][m], r[8][m].astype(np.int64)]) for r in res]
    return scal, np.stack(arrs)


# ---------------- deblocking (S/common/frame.c:621-792) ----------------
def make_deblock_info(g, seed, slice_b=0, psub8x8=1, cavlc_8x8dct=0, alpha=0, beta=0, chroma_off=0, chaos=False, qp_centre=30):
    """per-macroblock state in the reference's layouts (h->mb.type/qp/mb_transform_size/non_zero_count/ref/mv).
    chaos=False: what an encoder produces (vectors constant per partition, refs -1 for intra ...);
    chaos=True: every array independently random — the filter's decisions must match for ANY contents."""
    rng = np.random.default_rng(seed)
    W, H = g.mb_width, g.mb_height
    n = W * H
    if slice_b:
        types = np.array([0, 2, 7, 8, 12, 16, 17, 18], np.int8)
        tp = np.array([.05, .05, .1, .25, .15, .15, .1, .15])
    else:
        types = np.array([0, 1, 2, 4, 5, 6], np.int8)
        tp = np.array([.06, .04, .06, .44, .15, .25])
    mtype = rng.choice(types, n, p=tp).astype(np.int8)
    qp = np.clip(rng.normal(qp_centre, 7, n).round(), 0, 51).astype(np.int8)
    qp[rng.random(n) < 0.05] = rng.integers(0, 16)
    intra = mtype <= 3
    skip = (mtype == 6) | (mtype == 18)
    t8 = ((rng.random(n) < 0.3) & ~skip & (mtype != 2)).astype(np.int8)
    t8[mtype == 1] = 1
    nnz = np.zeros((n, 24), np.uint8)
    dense = rng.random(n) < 0.5
    nnz[:] = (rng.random((n, 24)) < np.where(dense, 0.5, 0.08)[:, None]) * rng.integers(1, 17, (n, 24))
    nnz[skip] = 0
    ref = [np.zeros((2 * H, 2 * W), np.int8) for _ in range(2)]
    mv = [np.zeros((4 * H, 4 * W, 2), np.int16) for _ in range(2)]
    base = np.array([-20, -12]) + rng.integers(-3, 4, 2)
    for mb in range(n):
        my, mx = divmod(mb, W)
        for l in range(2):
            if intra[mb] or (l == 1 and not slice_b):
                ref[l][2 * my:2 * my + 2, 2 * mx:2 * mx + 2] = -1
                continue
            r8 = rng.integers(0, 2, (2, 2)) if rng.random() < 0.15 else np.full((2, 2), rng.integers(0, 2) if rng.random() < 0.2 else 0)
            if slice_b and rng.random() < 0.3:
                r8 = np.full((2, 2), -1)  # list unused by this macroblock
            ref[l][2 * my:2 * my + 2, 2 * mx:2 * mx + 2] = r8
            v16 = base + rng.integers(-5, 6, 2)
            blk = np.broadcast_to(v16, (4, 4, 2)).copy()
            if mtype[mb] in (5, 17):  # 8x8 partitions, optionally 4x4 sub-partitions
                for by in range(2):
                    for bx in range(2):
                        blk[2 * by:2 * by + 2, 2 * bx:2 * bx + 2] = v16 + rng.integers(-4, 5, 2)
                        if psub8x8 and rng.random() < 0.4:
                            blk[2 * by:2 * by + 2, 2 * bx:2 * bx + 2] += rng.integers(-4, 5, (2, 2, 2))
            elif rng.random() < 0.3:  # 16x8 / 8x16
                if rng.random() < 0.5:
                    blk[2:] = v16 + rng.integers(-6, 7, 2)
                else:
                    blk[:, 2:] = v16 + rng.integers(-6, 7, 2)
            mv[l][4 * my:4 * my + 4, 4 * mx:4 * mx + 4] = blk
    if chaos:
        mtype = rng.integers(0, 19, n).astype(np.int8)
        t8 = rng.integers(0, 2, n).astype(np.int8)
        nnz = (rng.integers(0, 3, (n, 24)) * (rng.random((n, 24)) < 0.4)).astype(np.uint8)
        ref = [rng.integers(-1, 2, (2 * H, 2 * W)).astype(np.int8) for _ in range(2)]
        mv = [(base + rng.integers(-5, 6, (4 * H, 4 * W, 2))).astype(np.int16) for _ in range(2)]
    return {"alpha_c0_offset": alpha, "beta_offset": beta, "chroma_qp_offset": chroma_off, "b_slice_b": slice_b, "b_psub8x8": psub8x8,
            "b_cavlc_8x8dct": cavlc_8x8dct, "type": np.ascontiguousarray(mtype), "qp": np.ascontiguousarray(qp),
            "transform8x8": np.ascontiguousarray(t8), "nnz": np.ascontiguousarray(nnz),
            "ref0": np.ascontiguousarray(ref[0]), "ref1": np.ascontiguousarray(ref[1]),
            "mv0": np.ascontiguousarray(mv[0]), "mv1": np.ascontiguousarray(mv[1])}


def blocky_recon(clip, g, frame=0, seed=0):
    """a 'reconstructed' picture with coding artefacts: synthetic content quantised per 4x4/8x8 block (DC steps + noise)"""
    rng = np.random.default_rng(seed)
    y, u, v = clip.yuv420(frame)
    W16, H16 = 16 * g.mb_width, 16 * g.mb_height
    def pad(a, w, h):
        out = np.zeros((h, w), np.uint8)
        out[:a.shape[0], :a.shape[1]] = a
        out[a.shape[0]:, :a.shape[1]] = a[-1:, :]
        out[:, a.shape[1]:] = out[:, a.shape[1] - 1:a.shape[1]]
        return out
    y, u, v = pad(y, W16, H16), pad(u, W16 // 2, H16 // 2), pad(v, W16 // 2, H16 // 2)
    def blockify(a, bs):
        h, w = a.shape
        off = rng.integers(-6, 7, (h // bs, w // bs))
        return np.clip(a.astype(np.int32) + np.kron(off, np.ones((bs, bs), np.int32)) + rng.integers(-1, 2, a.shape), 0, 255).astype(np.uint8)
    return blockify(y, 4), blockify(u, 2), blockify(v, 2)


def padded_chroma(g, c):
    """(h/2, w/2) chroma picture -> plane with mod16 padding and 16-px replicated borders (x264_frame_expand_border, planes 1/2)"""
    H, W = g.lines // 2, g.mb_width * 8
    out = np.zeros((H + 32, W + 32), np.uint8)
    ys = np.clip(np.arange(-16, H + 16), 0, c.shape[0] - 1)
    xs = np.clip(np.arange(-16, W + 16), 0, c.shape[1] - 1)
    out[:] = c[np.ix_(ys, xs)]
    return out


def skip_probe_cases(seed, n):
    """macroblock (fenc, pred) tile pairs around the skip decision: pred = fenc + noise of a few amplitudes, flat offsets (DC-only
    differences), single-pixel spikes, identical tiles.  -> list of (qp, chroma_qp, fy, fu, fv, py, pu, pv)"""
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        fy = rng.integers(0, 256, (16, 16), dtype=np.uint8)
        fu = rng.integers(0, 256, (8, 8), dtype=np.uint8)
        fv = rng.integers(0, 256, (8, 8), dtype=np.uint8)
        kind = i % 6
        amp = (0, 1, 2, 3, 6, 12)[int(rng.integers(0, 6))]

        def near(a, amp):
            if kind == 0:
                return a.copy()
            if kind == 1:   # flat offset: only the DC differs
                return np.clip(a.astype(np.int32) + int(rng.integers(-amp, amp + 1)), 0, 255).astype(np.uint8)
            if kind == 2:   # one spike
                b = a.copy()
                b[int(rng.integers(0, a.shape[0])), int(rng.integers(0, a.shape[1]))] ^= int(rng.integers(0, 64))
                return b
            if kind == 3:   # luma clean, chroma noisy (the SSD gate and chroma DC/AC exits decide)
                amp2 = 0 if a.shape[0] == 16 else amp
                return np.clip(a.astype(np.int32) + rng.integers(-amp2, amp2 + 1, a.shape), 0, 255).astype(np.uint8)
            return np.clip(a.astype(np.int32) + rng.integers(-amp, amp + 1, a.shape), 0, 255).astype(np.uint8)
        py, pu, pv = near(fy, amp), near(fu, amp), near(fv, amp)
        qp = int(rng.integers(0, 52))
        cqp = int(rng.integers(0, 52)) if i % 3 else min(qp, 39)
        out.append((qp, cqp, fy, fu, fv, py, pu, pv))
    return out


def residual_digest(o, cases, cqm):
    """run the inter residual driver and the skip probe over skip_probe_cases -> (int32 summary rows, uint8 skip flags); the summary row
    of a case/flag pair is (cbp_luma, cbp_chroma, sum nnz, sum |coeff|, crc-ish weighted coefficient sum, sum of reconstructed pixels)"""
    import xo_api as X
    rows, skips = [], []
    for (qp, cqp, fy, fu, fv, py, pu, pv) in cases:
        for flags in range(4):
            r, ry, ru, rv = o.residual_inter_mb(X.ResidIn(qp, cqp, flags & 1, flags >> 1, cqm), fy, fu, fv, py, pu, pv)
            co = np.concatenate([np.array(r.luma8x8).reshape(-1) if flags & 1 else np.array(r.luma4x4)[:16].reshape(-1),
                                 np.array(r.luma4x4)[16:].reshape(-1), np.array(r.chroma_dc).reshape(-1)]).astype(np.int64)
            wsum = int((co * (np.arange(len(co)) % 251 + 1)).sum() & 0x7fffffff)
            rows.append((r.cbp_luma, r.cbp_chroma, int(np.array(r.nnz).sum()), int(np.abs(co).sum()), wsum,
                         int(ry.astype(np.int64).sum() + 3 * ru.astype(np.int64).sum() + 5 * rv.astype(np.int64).sum())))
        skips.append(o.probe_skip_mb(X.ResidIn(qp, cqp, 0, 1, cqm), fy, fu, fv, py, pu, pv))
    return np.array(rows, np.int32), np.array(skips, np.uint8)


def intra_cases(seed, n):
    """(neighbour mask, lambda, satd?, slice_b?, fenc tiles, neighbour vectors) for the Intra16x16 / chroma cost stage: smooth ramps (the
    plane predictor's home ground, incl. slopes that clip), flat, noise and 0/255 extremes; every neighbour-mask branch"""
    rng = np.random.default_rng(seed)
    out = []
    masks = [0xF, 0xB, 0x1, 0x2, 0x0, 0x3, 0x5, 0x6, 0xA, 0x9]
    for i in range(n):
        kind = i % 5
        yy, xx = np.mgrid[-1:16, -1:16]
        if kind == 0:
            big = rng.integers(0, 256, (17, 17))
        elif kind == 1:   # ramp + noise
            big = rng.integers(0, 256) + rng.integers(-20, 21) * xx + rng.integers(-20, 21) * yy + rng.integers(-3, 4, (17, 17))
        elif kind == 2:
            big = np.full((17, 17), int(rng.integers(0, 256))) + rng.integers(-2, 3, (17, 17))
        elif kind == 3:   # extremes
            big = rng.integers(0, 2, (17, 17)) * 255
        else:             # gentle ramp
            big = 128 + rng.integers(-4, 5) * xx + rng.integers(-4, 5) * yy + rng.integers(-1, 2, (17, 17))
        big = np.clip(big, 0, 255).astype(np.uint8)
        nby = np.concatenate([big[0, :1], big[0, 1:], big[1:, 0]]).astype(np.uint8)
        fy = np.ascontiguousarray(big[1:, 1:])
        if kind == 3 and i % 2:
            fy = (255 - fy).astype(np.uint8)
        chroma = []
        for c in range(2):
            b2 = np.clip(big[:9, :9].astype(np.int32) + rng.integers(-6, 7, (9, 9)), 0, 255).astype(np.uint8) if kind else \
                rng.integers(0, 256, (9, 9), dtype=np.uint8)
            chroma.append((np.ascontiguousarray(b2[1:, 1:]), np.concatenate([b2[0, :1], b2[0, 1:], b2[1:, 0]]).astype(np.uint8)))
        lam = int(rng.choice([1, 1, 2, 4, 6, 10, 18, 32, 57, 91]))
        out.append((masks[i % len(masks)] if i % 3 else 0xF, lam, int(rng.integers(0, 4) != 0), int(rng.integers(0, 2)), fy, chroma[0][0], chroma[1][0],
                    nby, chroma[0][1], chroma[1][1]))
    return out


def intra_digest(o, cases):
    """-> int32[n, 18]: cost16[7], cost_chroma[7], best16, best_chroma, mode16, mode_chroma of every case"""
    import xo_api as X
    rows = []
    for (nbr, lam, satd, sb, fy, fu, fv, nby, nbu, nbv) in cases:
        r = o.intra_mb_costs(X.IntraIn(nbr, lam, satd, sb), fy, fu, fv, nby, nbu, nbv).astuple()
        rows.append(list(r[0]) + list(r[1]) + list(r[2:]))
    return np.array(rows, np.int32)
