#!/usr/bin/env python3
"""SASS opcode histogram per kernel of the built library (cuobjdump -sass): what the kernels are made of, and the
absence / presence of the instruction families profiles/README.md talks about (DFMA vs DMUL+DADD, LDGSTS = cp.async,
UBLKCP / UTMA* = bulk copies, LDL / STL = local memory).  usage: sass_hist.py [LIB.so] > profiles/sass_r2.md"""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "odr_audioenc_b200", "libtoolame_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|\(.*$|^void ", "", name)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
watch = ["DFMA", "DMUL", "DADD", "DSETP", "MUFU", "F2I", "I2F", "LDS", "STS", "LDG", "STG", "LDGSTS", "LDL", "STL", "LDC", "SHFL",
         "BAR", "ATOMS", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "DMMA"]
print("| kernel | instructions | " + " | ".join(watch) + " | top opcodes |")
print("|---|---|" + "---|" * (len(watch) + 1))
for k, h in hist.items():
    fam = collections.Counter()
    for op, n in h.items():
        for w in watch:
            if op == w or op.startswith(w + "_") or (w in ("LDS", "STS", "LDG", "STG", "LDL", "STL") and op == w):
                fam[w] += n
    top = ", ".join("%s %d" % (op, n) for op, n in h.most_common(6))
    print("| %s | %d | " % (k, sum(h.values())) + " | ".join(str(fam.get(w, 0)) for w in watch) + " | %s |" % top)
print("\n`cuobjdump -sass` of %s; one sm_100a cubin.  DFMA appears only inside libdevice routines (log10, pow, sincos, "
      "atan2, division) and in the FP64 probe; the arithmetic of the path is DMUL / DADD.  LDGSTS = cp.async." % os.path.basename(lib))
