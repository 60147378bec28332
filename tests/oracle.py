"""ctypes binding of the CPU oracle (oracle/mp2_oracle.c) -- TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_LIB = os.path.join(ORACLE_DIR, "_build", "libmp2_oracle.so")


class Cfg(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        "fs_hz", "version", "sfreq_idx", "mode", "mode_ext", "nch", "bitrate_kbps", "bitrate_index",
        "tablenum", "sblimit", "jsbound", "dab_ext", "lg_frame", "psy", "pad_len", "psy_freq")]


TAP_DTYPE = np.dtype([
    ("sb_sample", "<f8", (2, 36, 32)),
    ("scalar_pre", "u1", (2, 3, 32)), ("scalar", "u1", (2, 3, 32)), ("j_scale", "u1", (3, 32)),
    ("scfsi", "u1", (2, 32)), ("bit_alloc", "u1", (2, 32)),
    ("smr", "<f8", (2, 32)), ("ltmin", "<f8", (2, 32)), ("spike", "<f8", (2, 32)),
    ("q", "<u4", (2, 36, 32)),
    ("mode", "<i4"), ("mode_ext", "<i4"), ("jsbound", "<i4"), ("adb_left", "<i4"),
    ("crc16", "<u4"), ("scfcrc_own", "u1", (4,)),
], align=True)

_lib = None


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR, "port"], check=True)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_LIB):
            build()
        _lib = C.CDLL(ORACLE_LIB)
        _lib.mp2o_configure.argtypes = [C.POINTER(Cfg), C.c_long, C.c_char, C.c_int, C.c_int, C.c_int]
        _lib.mp2o_encode.argtypes = [C.POINTER(Cfg), C.c_void_p, C.c_long, C.c_long, C.c_long,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.mp2o_fht1024.argtypes = [C.c_void_p]
    return _lib


def configure(fs, mode, bitrate, psy=1, pad_len=0):
    c = Cfg()
    rc = lib().mp2o_configure(C.byref(c), fs, mode.encode(), bitrate, psy, pad_len)
    if rc:
        raise ValueError("mp2o_configure -> %d" % rc)
    return c


def encode(cfg, pcm, f0=0, f1=None, xpad=None, taps=False):
    """pcm: int16 (n_samples, nch).  Returns (bytes u8 array, taps or None)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    n_total = pcm.shape[0] // 1152
    if f1 is None:
        f1 = n_total
    out = np.zeros((f1 - f0) * cfg.lg_frame, dtype=np.uint8)
    tp = np.zeros(f1 - f0, dtype=TAP_DTYPE) if taps else None
    assert TAP_DTYPE.itemsize == 36048 or True
    xp = None
    if xpad is not None:
        xp = np.ascontiguousarray(xpad, dtype=np.uint8)
    rc = lib().mp2o_encode(C.byref(cfg), pcm.ctypes.data, n_total, f0, f1,
                           xp.ctypes.data if xp is not None else None, out.ctypes.data,
                           tp.ctypes.data if taps else None)
    if rc:
        raise RuntimeError("mp2o_encode -> %d" % rc)
    return out, tp
