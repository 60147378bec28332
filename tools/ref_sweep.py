#!/usr/bin/env python3
"""Long-stream pin of the oracle port against the UNMODIFIED reference (build container, CPU only).

For every case a seeded signal of FRAMES frames is encoded twice: by oracle/_ref/ref_driver (the reference's own
libtoolame-dab compiled from /root/reference, one stateful process per stream) and by the stateless port
oracle/mp2_oracle.c in independent segments of 500 frames on all host cores.  The GPU sweeps
(tools/parity_sweep.py) compare the CUDA path with the port; this closes the triangle on the same kind of sample
size, so a rare path on which port and GPU agree with each other but not with the reference cannot hide.

usage: ref_sweep.py [FRAMES_PER_CASE=100000] [OUT.json=profiles/ref_sweep_r2.json] [CASE_FILTER]
"""
import json
import multiprocessing as mp
import os
import shutil
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import cases  # noqa: E402
import oracle  # noqa: E402
import reftool  # noqa: E402
from parity_sweep import make_long  # noqa: E402

# (config, signal, psy model, X-PAD): the BASELINE configurations A, B (s and j), C, E (psy 2), psy 0, X-PAD, plus
# the table-2 / mono / dual-channel / 32 kHz / 16 kHz corners
CASES = [("A", "S1", 1, False), ("Bs", "S8", 1, False), ("Bj", "S1", 1, False), ("Bj", "S8", 1, False),
         ("Bj", "S2", 1, False), ("Bj", "S4", 1, False), ("C", "S1", 1, False), ("C", "S8", 1, False),
         ("E1", "S1", 2, False), ("E1", "S8", 2, False), ("Bj", "S8", 0, False), ("Bj", "S1", 1, True),
         ("T2j", "S8", 1, False), ("M48", "S1", 1, False), ("D", "S8", 1, False), ("C", "S8", 2, True),
         ("R32", "S8", 1, False), ("R16", "S1", 1, False)]
SEG = 500


def _oracle_seg(job):
    cfg_name, psy, pad_len, f0, f1, pcm_path, xpad_path = job
    fs, mode, br = cases.CONFIGS[cfg_name]
    nch = 1 if mode == "m" else 2
    pcm = np.memmap(pcm_path, dtype=np.int16, mode="r").reshape(-1, nch)
    xpad = np.fromfile(xpad_path, dtype=np.uint8).reshape(-1, pad_len + 1) if xpad_path else None
    c = oracle.configure(fs, mode, br, psy, pad_len)
    out, _ = oracle.encode(c, pcm, f0, f1, xpad=xpad)
    return f0, out


def _ref_stream(job):
    cfg_name, psy, pad_len, pcm_path, xpad_path, out_path = job
    fs, mode, br = cases.CONFIGS[cfg_name]
    cmd = [reftool.REF_DRIVER, str(fs), mode, str(br), str(psy), str(pad_len), pcm_path, out_path]
    if xpad_path:
        cmd += ["--xpad", xpad_path]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        raise RuntimeError("ref_driver %s: %s" % (cfg_name, r.stderr[-500:]))
    return time.time() - t0


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    out_json = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "ref_sweep_r2.json")
    only = sys.argv[3] if len(sys.argv) > 3 else None
    if not reftool.have_ref():
        raise SystemExit("oracle/_ref is not built (needs /root/reference): make -C oracle ref")
    todo = [c for c in CASES if c[0] in cases.CONFIGS and (only is None or only in "%s-%s-psy%d" % c[:3])]
    tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    cores = os.cpu_count() or 1
    report, tot, tot_bad = [], 0, 0
    try:
        with mp.get_context("fork").Pool(cores) as pool:
            # in groups of `cores` cases: first the reference streams side by side (one process each), then the port
            for g0 in range(0, len(todo), cores):
                group = todo[g0:g0 + cores]
                jobs = []
                for k, (cfg_name, sig, psy, pad) in enumerate(group):
                    fs, mode, br = cases.CONFIGS[cfg_name]
                    nch = 1 if mode == "m" else 2
                    pcm = make_long(sig, n, nch, fs)
                    pcm_path = os.path.join(tmp, "pcm%d.bin" % k)
                    pcm.tofile(pcm_path)
                    del pcm
                    pad_len, xpad_path = 0, None
                    if pad:
                        pad_len = cases.PAD_LEN
                        xpad_path = os.path.join(tmp, "xpad%d.bin" % k)
                        cases.xpad_records(n, pad_len, seed=4242 + k).tofile(xpad_path)
                    jobs.append((cfg_name, psy, pad_len, pcm_path, xpad_path, os.path.join(tmp, "ref%d.mp2" % k)))
                ref_secs = pool.map(_ref_stream, jobs, chunksize=1)
                for (cfg_name, sig, psy, pad), job, rs in zip(group, jobs, ref_secs):
                    _, _, pad_len, pcm_path, xpad_path, ref_path = job
                    fs, mode, br = cases.CONFIGS[cfg_name]
                    lg = oracle.configure(fs, mode, br, psy, pad_len).lg_frame
                    ref = np.fromfile(ref_path, dtype=np.uint8)
                    assert ref.size == n * lg, (cfg_name, ref.size, n * lg)
                    ref = ref.reshape(n, lg)
                    t0 = time.time()
                    segs = [(cfg_name, psy, pad_len, f0, min(f0 + SEG, n), pcm_path, xpad_path) for f0 in range(0, n, SEG)]
                    bad = []
                    for f0, out in pool.imap_unordered(_oracle_seg, segs):
                        k = out.size // lg
                        d = np.flatnonzero((out.reshape(k, lg) != ref[f0:f0 + k]).any(axis=1))
                        bad += [f0 + int(x) for x in d]
                    rec = {"config": cfg_name, "sample_rate": fs, "mode": mode, "bitrate": br, "signal": sig, "psy": psy,
                           "xpad": bool(pad), "frames": n, "frames_differing": len(bad), "examples": sorted(bad)[:8],
                           "reference_seconds": round(rs, 1), "port_seconds": round(time.time() - t0, 1)}
                    print(json.dumps(rec), flush=True)
                    report.append(rec)
                    tot += n
                    tot_bad += len(bad)
                for f in os.listdir(tmp):
                    os.remove(os.path.join(tmp, f))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    summary = {"what": "oracle port (oracle/mp2_oracle.c, stateless 500-frame segments) vs the unmodified reference "
                       "(oracle/_ref/ref_driver, one stateful stream), output bytes per frame",
               "frames": tot, "frames_differing": tot_bad, "cases": report}
    print("TOTAL %d frames, %d differ" % (tot, tot_bad))
    json.dump(summary, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
