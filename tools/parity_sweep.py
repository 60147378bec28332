#!/usr/bin/env python3
"""Large-sample parity measurement (GPU box): encode long seeded signals with the CUDA path and with the oracle
(one CPU process per core), report the fraction of byte-identical frames and classify every differing frame by the
first stage at which it departs (SMR value -> bit allocation -> bytes).

usage: parity_sweep.py [FRAMES_PER_CASE] [OUT.json]"""
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import oracle  # noqa: E402
import signals  # noqa: E402

CASES = [("Bj", "S1"), ("Bj", "S2"), ("Bj", "S8"), ("Bj", "S6"), ("Bj", "S4"), ("Bj", "S7"), ("Bs", "S1"), ("Bs", "S8"),
         ("A", "S1"), ("A", "S8"), ("C", "S1"), ("C", "S8"), ("C", "S2"), ("T2j", "S8"), ("M48", "S1"), ("E1", "S8")]
PSY = int(os.environ.get("SWEEP_PSY", "1"))
if PSY == 2:
    CASES = [("E1", "S1"), ("E1", "S8"), ("E1", "S2"), ("E1", "S6"), ("E1", "S4"), ("Bs", "S8"), ("C", "S8"), ("A", "S1")]
SEG = 500  # frames per oracle work item


def _oracle_seg(job):
    cfg_name, sig, n, f0, f1, pcm_path = job
    fs, mode, br = cases.CONFIGS[cfg_name]
    nch = 1 if mode == "m" else 2
    pcm = np.memmap(pcm_path, dtype=np.int16, mode="r").reshape(-1, nch)  # written once by the parent
    c = oracle.configure(fs, mode, br, PSY)
    out, tap = oracle.encode(c, pcm, f0, f1, taps=True)
    return f0, out, tap["smr"].copy(), tap["bit_alloc"].copy(), tap["scalar"].copy()


def make_long(sig, n, nch, fs, piece=20000):
    """signals.make in pieces of `piece` frames (bounded memory); the pieces continue the same seeded signal where
    its definition is a function of absolute time / sample index, so up to `piece` frames it equals signals.make"""
    if n <= piece:
        return signals.make(sig, n, nch, fs)
    out = np.empty((n * 1152, nch), dtype=np.int16)
    for k, f0 in enumerate(range(0, n, piece)):
        m = min(piece, n - f0)
        if sig in ("S1", "S8", "S2"):
            out[f0 * 1152:(f0 + m) * 1152] = signals.SIGNALS[sig](m * 1152, nch, fs, seed={"S1": 12345, "S8": 777, "S2": 1}[sig] + 7919 * k)
        else:
            out[f0 * 1152:(f0 + m) * 1152] = signals.make(sig, m, nch, fs)
    return out


def main():
    import tempfile
    os.environ["TLB_HOST_CHUNK"] = str(1 << 30)  # one launch chunk per case, so that the taps cover every frame
    tmp_dir = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    out_path = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "parity_sweep.json")
    import odr_audioenc_b200 as tl
    report, tot, tot_bad = [], 0, 0
    with mp.get_context("fork").Pool(os.cpu_count()) as pool:
        for cfg_name, sig in CASES:
            fs, mode, br = cases.CONFIGS[cfg_name]
            nch = 1 if mode == "m" else 2
            t0 = time.time()
            pcm = make_long(sig, n, nch, fs)
            pcm_path = os.path.join(tmp_dir, "pcm.bin")
            pcm.tofile(pcm_path)
            enc = tl.BatchEncoder(fs, mode, br, PSY, chunk_frames=n + 1)
            got = enc.encode(pcm).reshape(n, -1)
            smr = enc.tap(tl.TAP_SMR, n)
            side = enc.tap(tl.TAP_SIDE, n)
            jobs = [(cfg_name, sig, n, f0, min(f0 + SEG, n), pcm_path) for f0 in range(0, n, SEG)]
            bad_frames, smr_off, first_stage = [], 0, {"scalefactor": 0, "bit_alloc": 0, "bytes_only": 0}
            for f0, want, o_smr, o_alloc, o_scalar in pool.imap_unordered(_oracle_seg, jobs):
                k = want.size // got.shape[1]
                w = want.reshape(k, -1)
                diff = np.flatnonzero((got[f0:f0 + k] != w).any(axis=1))
                smr_off += int((np.abs(smr[f0:f0 + k, :nch] - o_smr[:, :nch]) > 1e-9).any(axis=(1, 2)).sum())
                for d in diff:
                    f = f0 + int(d)
                    bad_frames.append(f)
                    if not np.array_equal(side["scalar"][f], o_scalar[d]):
                        first_stage["scalefactor"] += 1
                    elif not np.array_equal(side["bit_alloc"][f], o_alloc[d]):
                        first_stage["bit_alloc"] += 1
                    else:
                        first_stage["bytes_only"] += 1
            rec = {"config": cfg_name, "signal": sig, "psy": PSY, "frames": n, "frames_differing": len(bad_frames),
                   "frames_with_smr_off_by_1e-9": smr_off, "first_departure": first_stage,
                   "examples": sorted(bad_frames)[:8], "seconds": round(time.time() - t0, 1)}
            print(json.dumps(rec), flush=True)
            report.append(rec)
            tot += n
            tot_bad += len(bad_frames)
    summary = {"frames": tot, "frames_differing": tot_bad, "identical_fraction": 1 - tot_bad / tot, "cases": report}
    print("TOTAL %d frames, %d differ -> %.5f %% identical" % (tot, tot_bad, 100 * (1 - tot_bad / tot)))
    os.makedirs(os.path.dirname(out_path), exist_ok=True)
    json.dump(summary, open(out_path, "w"), indent=1)
    import shutil
    shutil.rmtree(tmp_dir, ignore_errors=True)


if __name__ == "__main__":
    main()
