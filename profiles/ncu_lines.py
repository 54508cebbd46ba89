#!/usr/bin/env python
"""Per-source-line instruction and stall-sample shares of one kernel from an .ncu-rep (SASS source page) joined
with nvdisasm's line info of the built object:  python profiles/ncu_lines.py rep.ncu-rep build/obj.o kernel_substr [top]"""
import csv, io, re, subprocess, sys, tempfile, os, collections

rep, obj, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
td = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, cubin)], capture_output=True, text=True).stdout
# walk the disassembly of the kernel's text section: "//## File "...", line N" markers precede instructions
lines_of = []   # per instruction (in order): source line
cur = None
inside = False
for ln in dis.splitlines():
    if ln.startswith(".text.") or ln.strip().startswith(".section"):
        inside = (kname in ln) and ".text." in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", ln) and ";" in ln:
        lines_of.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iI, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
if len(data) != len(lines_of):
    print("warning: %d SASS rows in the report, %d instructions in the object" % (len(data), len(lines_of)))
agg = collections.defaultdict(lambda: [0.0, 0.0])
for r, lo in zip(data, lines_of):
    agg[lo][0] += float(r[iI] or 0)
    agg[lo][1] += float(r[iS] or 0)
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
src = {}
print("total warp instructions %.4g, samples %d" % (ti, ts))
for lo, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    text = ""
    if lo:
        path = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", lo[0])
        if lo[0] not in src and os.path.exists(path):
            src[lo[0]] = open(path).read().splitlines()
        if lo[0] in src and lo[1] - 1 < len(src[lo[0]]):
            text = src[lo[0]][lo[1] - 1].strip()[:100]
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * v[0] / ti, 100 * v[1] / ts, lo[0] if lo else "?", lo[1] if lo else "?", text))
