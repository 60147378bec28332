#!/usr/bin/env python3
"""Per-source-line summary of an ncu report: joins `ncu --page source --csv` (SASS, with executed-instruction and
stall-sample counts) with the line table of the library's cubin (nvdisasm -g), instruction by instruction.

usage: ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTRING [LIB.so] [TOP_N]"""
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "odr_audioenc_b200", "libtoolame_b200.so")
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40

txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# several kernels may be in the report: take the first whose name matches
start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and ((kern + "(") in r[1] or (kern + "<") in r[1]))
hdr = rows[start + 1]
ci, cs, csrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
cti = hdr.index("Thread Instructions Executed")
sass = []
for r in rows[start + 2:]:
    if not r or r[0] == "Kernel Name":
        break
    sass.append((r[csrc].strip(), int(r[ci]), int(r[cs]), int(r[cti])))

with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    dis = ""
    for cub in glob.glob(os.path.join(td, "*.cubin")):
        dis += subprocess.run(["nvdisasm", "-g", "-c", cub], capture_output=True, text=True).stdout
sec = re.split(r"\n\s*\.section\s+\.text\.", dis)
body = next(s for s in sec[1:] if (kern + "E") in s.split("\n", 1)[0] or (kern + "I") in s.split("\n", 1)[0])
lines, cur = [], 0
for ln in body.split("\n"):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = int(m.group(2))
        continue
    m = re.match(r"\s+(/\*[0-9a-f]+\*/)?\s*([@!A-Z0-9_.]+[^;]*);", ln)
    if m and not ln.strip().startswith("."):
        lines.append(cur)
if len(lines) != len(sass):
    print("warning: %d instructions in the cubin vs %d in the report; aligning by index" % (len(lines), len(sass)))
agg = collections.defaultdict(lambda: [0, 0, 0])
for k, (s, n, smp, tn) in enumerate(sass):
    a = agg[lines[k] if k < len(lines) else -1]
    a[0] += n
    a[1] += smp
    a[2] += tn
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
src = open(os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", "mp2_kernels.cu")).read().split("\n")
print("kernel %s: %d warp instructions, %d stall samples" % (kern, tot_i, tot_s))
print(" inst%  smp%  thr/inst  line  source")
for line, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top_n]:
    text = src[line - 1].strip()[:100] if 0 < line <= len(src) else "?"
    print("%5.1f %5.1f %8.1f %5d  %s" % (100 * a[0] / tot_i, 100 * a[1] / tot_s, a[2] / max(a[0], 1), line, text))
