#!/usr/bin/env python3
"""compute-sanitizer workload (GPU box): a few short encodes across configurations, psy models and chunk boundaries,
each checked against the oracle.  Run as
    compute-sanitizer --tool memcheck  --error-exitcode 7 python tools/sanitize.py
    compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize.py
Round 1 result (profiles/README.md): memcheck 0 errors, racecheck 0 hazards."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402
import odr_audioenc_b200 as tl  # noqa: E402

ok = True
for cfg, sig, psy in (("Bj", "S8", 1), ("C", "S1", 1), ("E1", "S8", 2), ("T2", "S2", 1), ("L8", "S1", 1)):
    fs, mode, br, pcm, _, _ = cases.make_case(cfg, sig, 20)
    out = tl.BatchEncoder(fs, mode, br, psy, chunk_frames=9).encode(pcm)
    ref, _ = oracle.encode(oracle.configure(fs, mode, br, psy), pcm)
    same = bool((out == ref).all())
    ok &= same
    print(cfg, sig, "psy", psy, "identical" if same else "DIFFERENT")
sys.exit(0 if ok else 1)
