#!/usr/bin/env python
"""time_intra16.py — x264_cuda_residual_intra16 on every macroblock of one frame (raster job list = one wavefront launch): milliseconds per
call through the host-array entry point (job upload and coefficient read-back included) and for the device-resident form (CUDA events)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as ge  # noqa: E402

def main():
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    pkg = ge.load_pkg()
    from x264_vs2008_b200 import synth
    import torch
    ctx = pkg.Context(0)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    clip = synth.Clip(w, h, seed=5)
    y, u, v = clip.yuv420(0)
    fenc, fdec = ctx.frame(w, h, pkg.FRAME_CHROMA), ctx.frame(w, h, pkg.FRAME_CHROMA)
    fenc.upload(y); fenc.upload_chroma(u, v); fenc.expand_border_mod16()
    fdec.upload(y); fdec.upload_chroma(u, v); fdec.expand_border_mod16()
    ctx.set_quant_preset(0)
    mbw, mbh = (w + 15) // 16, (h + 15) // 16
    jobs = np.zeros(mbw * mbh, pkg.INTRA16_JOB)
    rng = np.random.default_rng(1)
    for i in range(len(jobs)):
        mx, my = i % mbw, i // mbw
        jobs[i]["mb_x"], jobs[i]["mb_y"], jobs[i]["qp"], jobs[i]["chroma_qp"] = mx, my, 26, 26
        m = int(rng.integers(0, 4)) if mx and my else (1 if mx else 0 if my else 6)
        jobs[i]["mode16"] = m if (mx and my) else (4 if mx and not my else 5 if my and not mx else 6)
        jobs[i]["mode_chroma"] = int(rng.integers(0, 4)) if mx and my else (4 if mx and not my else 5 if my and not mx else 6)
    ts = []
    for _ in range(6):
        t = time.perf_counter(); ctx.residual_intra16(fenc, fdec, jobs); ts.append((time.perf_counter() - t) * 1e3)
    diag = jobs[np.argsort(jobs["mb_x"].astype(np.int32) + jobs["mb_y"], kind="stable")]  # the _dev entry takes tickets in list order
    d_jobs = torch.from_numpy(diag.view(np.uint8).reshape(-1)).cuda()
    d_out = torch.empty(len(jobs) * pkg.MB_COEFFS_I16.itemsize, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    dev = []
    for _ in range(6):
        ev[0].record(stream)
        ctx.check(pkg.lib().x264_cuda_residual_intra16_dev(ctx.h, fenc.h, fdec.h, d_jobs.data_ptr(), len(jobs), d_out.data_ptr()))
        ev[1].record(stream); ev[1].synchronize()
        dev.append(ev[0].elapsed_time(ev[1]))
    print("intra16 %dx%d: %d macroblocks, host-array call %.3f ms (min of 5), device-resident %.3f ms (min of 5)" % (w, h, len(jobs), min(ts[1:]), min(dev[1:])))

if __name__ == "__main__":
    main()
