#!/usr/bin/env python
"""encode_fps.py — encode one synthetic clip with the reference CLI (oracle/_ref/x264) and with the performance-mode build
(integration/_build/x264_b200), compare the streams byte for byte and print one JSON line with both frame rates."""
import argparse, json, os, re, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from test_integration_host import _clip, REF
B200 = os.path.join(ROOT, "integration", "_build", "x264_b200")


def run(exe, opts, src, out, w, h, threads=1, env=None):
    cmd = [exe, "--no-asm", "--threads", str(threads)] + opts.split() + ["-o", out, src, "%dx%d" % (w, h)]
    e = dict(os.environ); e.update(env or {})
    t = time.perf_counter()
    r = subprocess.run(cmd, capture_output=True, text=True, env=e)
    wall = time.perf_counter() - t
    m = re.search(r"encoded (\d+) frames, ([0-9.]+) fps", r.stderr)
    return {"rc": r.returncode, "wall_s": wall, "frames": int(m.group(1)) if m else 0, "fps": float(m.group(2)) if m else 0.0, "stderr": r.stderr}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--frames", type=int, default=12)
    ap.add_argument("--opts", default="--qp 26 --me esa --merange 16 --subme 2 --no-psnr --no-ssim")
    ap.add_argument("--ref-threads", type=int, default=0, help="also time the reference with this many frame threads (0: skip)")
    ap.add_argument("--tmp", default="/tmp")
    a = ap.parse_args()
    w, h = map(int, a.size.split("x"))
    src = os.path.join(a.tmp, "enc_%dx%d_%d.yuv" % (w, h, a.frames))
    if not os.path.exists(src):
        _clip(w, h, a.frames, src)
    o0, o1 = os.path.join(a.tmp, "ref.264"), os.path.join(a.tmp, "b200.264")
    ours = run(B200, a.opts, src, o1, w, h, env={"X264_B200_VERBOSE": "1"})
    ref = run(REF, a.opts, src, o0, w, h)
    line = {"size": a.size, "frames": a.frames, "opts": a.opts, "fps_b200": ours["fps"], "fps_ref_1thread": ref["fps"],
            "identical": ours["rc"] == 0 and ref["rc"] == 0 and open(o0, "rb").read() == open(o1, "rb").read(),
            "b200_report": [l for l in ours["stderr"].splitlines() if l.startswith("x264_b200")]}
    if a.ref_threads:
        rt = run(REF, a.opts, src, os.path.join(a.tmp, "reft.264"), w, h, threads=a.ref_threads)
        line["fps_ref_%dthreads" % a.ref_threads] = rt["fps"]
    if ours["rc"]:
        line["stderr"] = ours["stderr"][-1500:]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
