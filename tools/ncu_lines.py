#!/usr/bin/env python
"""per-source-line samples / instructions of one kernel: joins an `ncu --import-source on` report (source page, SASS order) with the
line table of the object file (nvdisasm -g).  usage: ncu_lines.py report.ncu-rep object.o kernel_substring source.cu [top_n]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, obj, kname, srcfile = sys.argv[1:5]
top_n = int(sys.argv[5]) if len(sys.argv) > 5 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(dis) if kname in l and l.startswith(".text.")][0]
cur, lines = None, []
base = os.path.basename(srcfile)
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        if lines: break
    m = re.search(r'//## File "(.*?)", line (\d+)(.*)', l)
    if m:
        chain = [(os.path.basename(m.group(1)), int(m.group(2)))] + [(os.path.basename(a), int(b)) for a, b in re.findall(r'inlined at "(.*?)", line (\d+)', m.group(3))]
        cur = chain
        continue
    if re.match(r'\s+/\*[0-9a-f]{4,5}\*/\s', l):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, R = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
assert len(R) == len(lines), (len(R), len(lines))
samp, ins = collections.Counter(), collections.Counter()
for k, r in enumerate(R):
    ch = lines[k] or [("?", 0)]
    key = next((c for c in ch if c[0] == base), ch[-1])
    samp[key] += float(r[ix["# Samples"]] or 0); ins[key] += float(r[ix["Instructions Executed"]] or 0)
tot, toti = sum(samp.values()), sum(ins.values())
src = open(srcfile).read().splitlines()
print("total samples %d, warp instructions %.2fM" % (tot, toti / 1e6))
for key, s in sorted(ins.items(), key=lambda x: -x[1])[:top_n]:
    f, ln = key
    t = src[ln - 1].strip()[:110] if f == base and 0 < ln <= len(src) else f
    print("%5s instr %5.1f%% samp %5.1f%%  %s" % (ln, 100 * ins[key] / toti, 100 * samp[key] / tot, t))
