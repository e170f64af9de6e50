"""Stall samples of one kernel of an .ncu-rep by CUDA source line: the SASS rows of `ncu --page source` are matched, in
order, with the instructions of `nvdisasm -g` (which carries //## File/line markers) of the same kernel in the object file.
usage: python tools/ncu_lines.py rep.ncu-rep kernel_regex object.o mangled_substring [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os

rep, kre, obj, mangled = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 30
sel = ["--kernel-id", kre[3:]] if kre.startswith("id:") else ["--kernel-name", "regex:" + kre]      # id:::::N = N-th launch
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# first kernel only
start = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blk = rows[start[0] + 1:(start[1] if len(start) > 1 else len(rows))]
hdr = blk[0]
si, ii = hdr.index("# Samples"), hdr.index("Instructions Executed")
sass = [(r[1].strip(), int(r[si] or 0), int(r[ii] or 0)) for r in blk[1:] if len(r) > si]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout.splitlines()
# locate the function
lines, cur, infn = [], None, False
for l in dis:
    if l.startswith(".text.") or re.match(r"^\s*\.section\s+\.text\.", l):
        infn = mangled in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(cur)
print(f"{len(sass)} SASS rows in the report, {len(lines)} instructions in the object", file=sys.stderr)
n = min(len(sass), len(lines))
agg, execs = {}, {}
for (txt, s, e), ln in zip(sass[:n], lines[:n]):
    agg[ln] = agg.get(ln, 0) + s
    execs[ln] = execs.get(ln, 0) + e
tot = sum(agg.values())
src = {}
for ln in sorted(agg, key=lambda k: -agg[k])[:top]:
    if ln is None:
        print(f"{agg[ln]:6d} {100*agg[ln]/tot:5.1f}%  inst {execs[ln]:9d}  (no line)")
        continue
    f, no = ln
    if f not in src:
        p = os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "csrc", f)
        src[f] = open(p).read().splitlines() if os.path.exists(p) else []
    text = src[f][no - 1].strip() if no - 1 < len(src[f]) else ""
    print(f"{agg[ln]:6d} {100*agg[ln]/tot:5.1f}%  inst {execs[ln]:9d}  {f}:{no}  {text[:110]}")
