"""The drop-in at the C level: a driver written against <toolame.h> compiles with the REFERENCE's own header (when
/root/reference is present) and with include/toolame.h, and links against libtoolame_b200.so; on the GPU it then
produces the oracle's bytes."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_INC = "/root/reference/libtoolame-dab"
LIBDIR = os.path.join(ROOT, "odr_audioenc_b200")
SRC = os.path.join(ROOT, "tests", "dropin_driver.c")


def _build(inc, exe):
    subprocess.run(["gcc", "-O2", "-std=c99", "-Wall", "-Werror", "-I" + inc, "-o", exe, SRC, "-L" + LIBDIR, "-ltoolame_b200",
                    "-Wl,-rpath," + LIBDIR], check=True)


def test_links_with_our_header(tmp_path):
    _build(os.path.join(ROOT, "include"), str(tmp_path / "drv"))


@pytest.mark.skipif(not os.path.isdir(REF_INC), reason="/root/reference absent")
def test_links_with_the_reference_header(tmp_path):
    _build(REF_INC, str(tmp_path / "drv"))
    exported = subprocess.run(["nm", "-D", "--defined-only", os.path.join(LIBDIR, "libtoolame_b200.so")],
                              capture_output=True, text=True, check=True).stdout
    for sym in open("/root/reference/libtoolame-dab.sym").read().split():
        assert (" T " + sym + "\n") in exported, sym


@pytest.mark.gpu
@pytest.mark.parametrize("fs,mode,br,psy", [(48000, "j", 128, 1), (24000, "m", 64, 1), (48000, "s", 192, 2)])
def test_driver_output_equals_oracle(tmp_path, fs, mode, br, psy):
    import oracle
    import signals
    exe = str(tmp_path / "drv")
    _build(os.path.join(ROOT, "include"), exe)
    nch = 1 if mode == "m" else 2
    n = 40
    pcm = signals.make("S8", n, nch, fs)
    pcm.tofile(tmp_path / "in.pcm")
    subprocess.run([exe, str(fs), mode, str(br), str(psy), str(tmp_path / "in.pcm"), str(tmp_path / "out.mp2")], check=True)
    want, _ = oracle.encode(oracle.configure(fs, mode, br, psy), pcm)
    assert np.array_equal(np.fromfile(tmp_path / "out.mp2", dtype=np.uint8), want)
