#!/usr/bin/env python
"""gops_scaling.py — frame rate of the GOP-parallel front end for 1, 2, 4, ... encoder threads on one GPU, with the per-thread reports"""
import json, os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
src = bench._enc_clip(1, int(sys.argv[1]) if len(sys.argv) > 1 else 288)
out = []
for w in [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else "1,2,4,8,12")]:
    n = min(w * bench.ENC_KEYINT * (int(sys.argv[3]) if len(sys.argv) > 3 else 1), int(sys.argv[1]) if len(sys.argv) > 1 else 288)
    cmd = [bench.B200_GOPS, "--no-asm"] + bench.ENC_OPTS.split() + ["--keyint", str(bench.ENC_KEYINT), "--workers", str(w), "--frames", str(n), "-o", "/tmp/g.264", src, "%dx%d" % (bench.W, bench.H)]
    e = dict(os.environ); e["X264_B200_VERBOSE"] = "1"
    t = time.perf_counter(); r = subprocess.run(cmd, capture_output=True, text=True, env=e); wall = time.perf_counter() - t
    m = re.search(r"encoded (\d+) frames, ([0-9.]+) fps", r.stderr)
    rep = [l for l in r.stderr.splitlines() if "host time in device calls" in l]
    out.append({"workers": w, "frames": n, "fps_wall": n / wall, "fps_inner": float(m.group(2)) if m else 0, "reports": rep[:3]})
    print(json.dumps(out[-1]), flush=True)
