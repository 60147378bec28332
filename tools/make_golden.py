#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref (make -C oracle ref).

Each fixture holds the reference's output bytes and decision taps (scalefactor indices, scfsi, bit allocation,
mode_ext, SMR) for a seeded synthetic signal (tests/signals.py) -- the PCM itself is regenerated from the seed.
Run in the build container only (needs /root/reference); the fixtures travel with the repo.  Existing fixtures are
kept (npz files are not byte-reproducible); pass --force to regenerate them all."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import reftool  # noqa: E402

out = os.path.join(ROOT, "tests", "golden")
os.makedirs(out, exist_ok=True)
for psy, cfg, sig, n in [(1,) + g for g in cases.GOLDEN] + [(2,) + g for g in cases.GOLDEN_PSY2] + [(0,) + g for g in cases.GOLDEN_PSY0]:
    path = os.path.join(out, ("%s_%s.npz" if psy == 1 else "psy%d_%%s_%%s.npz" % psy) % (cfg, sig))
    if os.path.exists(path) and "--force" not in sys.argv:
        continue
    fs, mode, br, pcm, pad_len, xpad = cases.make_case(cfg, sig, n)
    r = reftool.run_ref(pcm, fs, mode, br, psy, pad_len, xpad=xpad, taps=True, tapbig=True)
    t = r["tap"]
    np.savez_compressed(
        path, bytes=r["bytes"],
        pcm_crc=np.uint32(np.bitwise_xor.reduce(pcm.astype(np.uint16).ravel().astype(np.uint32) * np.arange(1, pcm.size + 1, dtype=np.uint32))),
        scalar=t["scalar"].astype(np.uint8), j_scale=t["j_scale"].astype(np.uint8), scfsi=t["scfsi"].astype(np.uint8),
        bit_alloc=t["bit_alloc"].astype(np.uint8), mode=t["mode"], mode_ext=t["mode_ext"], jsbound=t["jsbound"],
        smr=t["smr"], sb_first=r["big"]["sb_sample"][0], q_first=r["big"]["subband"][0])
    print(cfg, sig, len(r["bytes"]))
