#!/usr/bin/env python3
"""Warp-stall samples of one kernel per CUDA source line: joins the SASS page of an ncu report
(ncu -i REP --page source --csv --kernel-name regex:NAME) with the line table of the cubin (nvdisasm -g).

usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX LIBRARY.so [TOP=40]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, kern, so = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                          capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    hi = next(i for i, r in enumerate(rows) if "Source" in r and "Address" in r)
    full_name = next(r[1] for r in rows if r and r[0] == "Kernel Name")   # demangled, template arguments included
    hdr = rows[hi]
    ai, si, ni = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Source"), hdr.index("Instructions Executed")
    sass = []
    for r in rows[hi + 1:]:
        if len(r) <= ai or not r[0].startswith("0x"):
            break  # the first kernel instance only
        sass.append((int(r[0], 16), r[si].strip(), int(r[ai]), int(r[ni])))
    base = sass[0][0]
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.startswith("mp2_kernels.")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    line_of, cur, inside = {}, None, False
    demangled = {}
    for ln in dis.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            sym = ln.split(".text.")[1].split()[0]
            if sym not in demangled:   # the instance of a template that the report holds, not its siblings
                demangled[sym] = subprocess.run(["cu++filt", sym], capture_output=True, text=True).stdout.strip()
            inside = "$" not in sym and demangled[sym].replace("void ", "") == full_name.replace("void ", "")
        if not inside:
            continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:   # lines of other files (CUDA's headers: intrinsics, atomics) are shown under the header's name
            cur = int(m.group(2)) if m.group(1) == "mp2_kernels.cu" else "%s:%s" % (m.group(1), m.group(2))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    per_line = {}
    total = sum(s for _, _, s, _ in sass)
    for addr, text, s, n in sass:
        L = line_of.get(addr - base)
        a = per_line.setdefault(L, [0, 0])
        a[0] += s
        a[1] += n
    src = open(os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", "mp2_kernels.cu")).read().splitlines()
    print("%s: %d samples, %d warp instructions" % (kern, total, sum(n for _, _, _, n in sass)))
    for L, (s, n) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[L - 1].strip()[:100] if isinstance(L, int) and L <= len(src) else "(%s)" % (L or "no line")
        print("%5s %6.2f%% %10d  %s" % (L if isinstance(L, int) else "", 100.0 * s / total, n, text))


if __name__ == "__main__":
    main()
