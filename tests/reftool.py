"""Run the compiled reference (oracle/_ref/ref_driver) on a PCM array -- TEST INFRASTRUCTURE.

One subprocess per stream (the reference keeps its state in statics).
"""
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_DRIVER = os.path.join(REF_DIR, "ref_driver")
REF_LIB = os.path.join(REF_DIR, "libtoolame_ref.so")

TAP_DTYPE = np.dtype([
    ("mode", "<i4"), ("mode_ext", "<i4"), ("jsbound", "<i4"), ("sblimit", "<i4"),
    ("nch", "<i4"), ("tablenum", "<i4"), ("bitrate_index", "<i4"), ("dab_extension", "<i4"),
    ("scalar", "<u4", (2, 3, 32)), ("j_scale", "<u4", (3, 32)),
    ("scfsi", "<u4", (2, 32)), ("bit_alloc", "<u4", (2, 32)),
    ("smr", "<f8", (2, 32)), ("max_sc", "<f8", (2, 32)),
])
TAPBIG_DTYPE = np.dtype([("sb_sample", "<f8", (2, 3, 12, 32)), ("subband", "<u4", (2, 3, 12, 32))])


def have_ref():
    return os.path.exists(REF_DRIVER)


def run_ref(pcm, fs, mode, bitrate, psy=1, padlen=0, xpad=None, taps=False, tapbig=False):
    """pcm: int16 (n_samples, nch) interleaved.  Returns dict(bytes=..., tap=..., big=...)."""
    nch = 1 if mode == "m" else 2
    assert pcm.dtype == np.int16 and pcm.ndim == 2 and pcm.shape[1] == nch
    with tempfile.TemporaryDirectory() as td:
        pin = os.path.join(td, "in.pcm")
        pout = os.path.join(td, "out.mp2")
        np.ascontiguousarray(pcm).tofile(pin)
        cmd = [REF_DRIVER, str(fs), mode, str(bitrate), str(psy), str(padlen), pin, pout]
        if xpad is not None:
            px = os.path.join(td, "xpad.bin")
            np.ascontiguousarray(xpad, dtype=np.uint8).tofile(px)
            cmd += ["--xpad", px]
        if taps:
            cmd += ["--tap", os.path.join(td, "tap.bin")]
        if tapbig:
            cmd += ["--tapbig", os.path.join(td, "big.bin")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_driver failed (%d): %s" % (r.returncode, r.stderr[-2000:]))
        out = {"bytes": np.fromfile(pout, dtype=np.uint8), "stderr": r.stderr}
        if taps:
            out["tap"] = np.fromfile(os.path.join(td, "tap.bin"), dtype=TAP_DTYPE)
        if tapbig:
            out["big"] = np.fromfile(os.path.join(td, "big.bin"), dtype=TAPBIG_DTYPE)
        return out
