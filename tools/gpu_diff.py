#!/usr/bin/env python3
"""Diagnostic (GPU box): where does the CUDA path differ from the oracle?  usage: gpu_diff.py CFG SIG [N]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402
import odr_audioenc_b200 as tl  # noqa: E402

cfg, sig = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 40
fs, mode, br, pcm, pad_len, xpad = cases.make_case(cfg, sig, n)
c = oracle.configure(fs, mode, br, 1, pad_len)
ref, tap = oracle.encode(c, pcm, xpad=xpad, taps=True)
e = tl.BatchEncoder(fs, mode, br, 1, pad_len)
out = e.encode(pcm, xpad=xpad)
nch, sbl = c.nch, c.sblimit
bad = np.flatnonzero((out.reshape(n, -1) != ref.reshape(n, -1)).any(axis=1))
print("frames differing:", bad.size, "of", n, bad[:20])
smr = e.tap(tl.TAP_SMR, n)[:, :nch, :sbl]
want = tap["smr"][:, :nch, :sbl]
d = np.abs(smr - want)
print("smr max abs diff", d.max(), "n>1e-9:", int((d > 1e-9).sum()), "n!=0:", int((d != 0).sum()), "of", d.size)
for f, ch, sb in list(zip(*np.nonzero(d > 1e-9)))[:20]:
    print("  frame %d ch %d sb %d: gpu %.17g oracle %.17g" % (f, ch, sb, smr[f, ch, sb], want[f, ch, sb]))
side = e.tap(tl.TAP_SIDE, n)
for k in ("scfsi", "bit_alloc"):
    print(k, "mismatches:", int((side[k][:, :nch, :sbl] != tap[k][:, :nch, :sbl]).sum()))
sb_s = e.tap(tl.TAP_SB_SAMPLE, n)
print("sb_sample bit-identical:", bool(np.array_equal(sb_s, tap["sb_sample"][:, :nch])))
