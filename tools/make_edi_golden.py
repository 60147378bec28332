#!/usr/bin/env python3
"""Generate tests/golden/edi_*.npz with the reference's own EDI packetiser (oracle/_ref/edi_ref_driver: the
unmodified contrib/edioutput sources behind the call sequence of Output::EDI::write_frame).  Build container only."""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import edi_cases  # noqa: E402

DRIVER = os.path.join(ROOT, "oracle", "_ref", "edi_ref_driver")


def run_ref(case, frames, peaks, fec=None, chunk_len=None):
    """AF packets of the reference's packetiser; with fec (>= 0) also the PF fragments of its PFT layer per packet:
    returns [af, ...] or [(af, [fragment, ...]), ...]"""
    n, frame_len = frames.shape
    with tempfile.TemporaryDirectory() as td:
        rec = np.zeros((n, frame_len + 4), dtype=np.uint8)
        rec[:, :frame_len] = frames
        rec[:, frame_len:] = peaks.astype(np.int16).view(np.uint8).reshape(n, 4)
        rec.tofile(os.path.join(td, "in.bin"))
        extra = [] if fec is None else [str(fec)] + ([str(chunk_len)] if chunk_len else [])
        subprocess.run([DRIVER, str(int(case["tist"])), str(case["delay_ms"]), str(case["alignment"]), str(case["tai"]),
                        str(case["start"]), case["tag"], str(frame_len), os.path.join(td, "in.bin"), os.path.join(td, "out.bin")] + extra,
                       check=True)
        raw = open(os.path.join(td, "out.bin"), "rb").read()
    out, at = [], 0

    def take():
        nonlocal at
        size = int(np.frombuffer(raw[at:at + 4], dtype=np.uint32)[0])
        at += 4 + size
        return raw[at - size:at]

    while at < len(raw):
        af = take()
        if fec is None:
            out.append(af)
        else:
            count = int(np.frombuffer(raw[at:at + 4], dtype=np.uint32)[0])
            at += 4
            out.append((af, [take() for _ in range(count)]))
    return out


if __name__ == "__main__":
    for name, case in edi_cases.CASES.items():
        frames, peaks = edi_cases.inputs(case)
        pk = run_ref(case, frames, peaks)
        sizes = np.array([len(p) for p in pk], dtype=np.uint32)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "edi_%s.npz" % name), sizes=sizes,
                            data=np.frombuffer(b"".join(pk), dtype=np.uint8))
        print(name, len(pk), "packets", int(sizes.sum()), "bytes")
    # PFT layer: fragments of the first packets of two cases, fragmentation only and with Reed-Solomon protection
    for name, fec, chunk_len, count in edi_cases.PFT_CASES:
        case = dict(edi_cases.CASES[name], n=count)
        frames, peaks = edi_cases.inputs(edi_cases.CASES[name])
        res = run_ref(case, frames[:count], peaks[:count], fec, chunk_len)
        frags = [f for _, fl in res for f in fl]
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pft_%s_m%d_k%d.npz" % (name, fec, chunk_len)),
                            per_packet=np.array([len(fl) for _, fl in res], dtype=np.uint32),
                            sizes=np.array([len(f) for f in frags], dtype=np.uint32), data=np.frombuffer(b"".join(frags), dtype=np.uint8))
        print("pft", name, fec, chunk_len, len(frags), "fragments")
