#!/usr/bin/env python
"""One 1080p frame through every frame-batched kernel (for ncu launch lists / per-kernel profiles and device timing):
upload -> border -> hpel+integral -> lowres -> MB-batched ESA -> sub-pel refine -> MC -> inter residual -> deblock -> lowres frame cost."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as ge  # noqa: E402
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--time", action="store_true", help="print per-stage device times (CUDA events via torch)")
    a = ap.parse_args()
    pkg = ge.load_pkg()
    from x264_vs2008_b200 import synth
    import torch
    ctx = pkg.Context(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    w, h = a.width, a.height
    clip = synth.Clip(w, h, seed=3)
    flags = pkg.FRAME_HPEL | pkg.FRAME_INTEGRAL | pkg.FRAME_LOWRES | pkg.FRAME_CHROMA
    fenc, fref, fdec = ctx.frame(w, h, flags), ctx.frame(w, h, flags), ctx.frame(w, h, flags)
    y1, u1, v1 = clip.yuv420(1)
    y0, u0, v0 = clip.yuv420(0)
    g = fenc.g
    jobs = bench.build_jobs(pkg, g.mb_width, g.mb_height)
    mbjobs = bench.to_mb_jobs(pkg, jobs, g.mb_width, g.mb_height)
    ctx.set_cost_mv(bench.QP); ctx.set_quant_preset(0)
    rj = np.zeros(g.mb_width * g.mb_height, pkg.RESID_JOB)
    rj["mb_x"], rj["mb_y"] = np.tile(np.arange(g.mb_width), g.mb_height), np.repeat(np.arange(g.mb_height), g.mb_width)
    rj["qp"], rj["chroma_qp"], rj["flags"] = 26, 26, pkg.RESID_DECIMATE
    from helpers import make_deblock_info
    dinfo = make_deblock_info(g, seed=7)
    fref.init_lowres()
    for f in (fenc, fref):
        f.lookahead_alloc(2)
    stages = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        stages.setdefault(name, []).append(e0.elapsed_time(e1))

    for rep in range(a.reps):
        fenc.upload(y1); fenc.upload_chroma(u1, v1); fref.upload(y0); fref.upload_chroma(u0, v0)
        timed("border x2", lambda: (fenc.expand_border(), fref.expand_border()))
        timed("hpel+integral", fref.filter)
        timed("lowres", fenc.init_lowres)
        res = [None]
        timed("me_search_mb (73440 searches)", lambda: res.__setitem__(0, ctx.me_search_mb(fenc, fref, bench.ME_RANGE, mbjobs)))
        r = res[0]["part"].reshape(-1)
        j2 = jobs.copy()
        j2["seed_mv"][:, 0], j2["seed_mv"][:, 1], j2["seed_cost"] = r["bmx"], r["bmy"], r["bcost"]
        j2["mv_min_spel"] = (j2["mv_min_fpel"].astype(np.int32) - 5) * 4
        j2["mv_max_spel"] = (j2["mv_max_fpel"].astype(np.int32) + 5) * 4
        j2["flags"] = pkg.ME_MBCMP_SATD
        fin = [None]
        timed("subpel refine subme4 (73440 searches)", lambda: fin.__setitem__(0, ctx.me_search_small(fenc, fref, pkg.ME_METHOD_SEEDED, bench.ME_RANGE, 4, j2)))
        f16 = fin[0][0::9]
        mc = np.zeros(len(f16), pkg.MC_JOB)
        mc["bx"], mc["by"], mc["mvx"], mc["mvy"], mc["w"], mc["h"] = jobs["bx"][0::9], jobs["by"][0::9], f16["mv"][:, 0], f16["mv"][:, 1], 16, 16
        timed("mc_blocks (8160 MB)", lambda: ctx.mc_blocks(fref, fdec, mc))
        timed("residual_inter (8160 MB)", lambda: ctx.residual_inter(fenc, fdec, rj))
        timed("deblock (wavefront, P slice)", lambda: ctx.frame_deblock(fdec, dinfo))
        timed("lowres intra+P cost (x264_slicetype_frame_cost b=p1=1)", lambda: ctx.lowres_frame_cost(fenc, fref, fenc, 0, 1, 1, do_search=(1, 0)))
    ctx.synchronize()
    if a.time:
        for k, v in stages.items():
            print("%-42s %.3f ms (min of %d, includes H2D/D2H of job+result arrays where the entry point takes host arrays)" % (k, min(v), len(v)))
    print("launches:", ctx.launches())


if __name__ == "__main__":
    main()
