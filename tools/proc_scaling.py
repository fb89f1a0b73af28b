#!/usr/bin/env python
"""proc_scaling.py — the device-assisted encoder as N worker PROCESSES (one encoder, one CUDA context each; x264-vs2008_b200/gop_shard.py)
instead of N threads in one process (integration/x264_b200_gops.c): frames per second over the same clip."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import __graft_entry__ as ge
ge.load_pkg()
from x264_vs2008_b200 import gop_shard as G
n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1152
src = bench._enc_clip(1, n_frames)
for w in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "12,16")]:
    runs = G.split_runs(G.plan_gops(n_frames, bench.ENC_KEYINT), w)
    parts, wall = G.encode_gops(bench.B200_CLI, src, bench.W, bench.H, bench.ENC_OPTS.split(), bench.ENC_KEYINT, runs, "/tmp/proc_scaling", workers=w,
                                env={"X264_B200_DEVICE": "0"})
    print(json.dumps({"worker_processes": w, "frames": n_frames, "fps_wall": n_frames / wall}), flush=True)
