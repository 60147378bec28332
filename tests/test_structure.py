"""CPU: the independent structural checker (tests/mp2_check.py) accepts what the reference produced (golden
fixtures) and rejects corrupted streams -- so that its verdict on the GPU output (tests/test_gpu_parity.py) means
something."""
import os

import numpy as np
import pytest

import cases
import mp2_check

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["Bj_S1", "A_S8", "C_S2", "T2_S1", "T2j_S8", "M48_S1", "D_S8", "E1_S2", "L2_S1", "Bj_S3",
                                  "Bj_PAD", "psy2_E1_S8", "psy0_C_S1"])
def test_checker_accepts_reference_streams(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    frames = mp2_check.check_stream(g["bytes"])
    assert len(frames) == 10
    assert np.array_equal(np.array([f["mode_ext"] for f in frames]), g["mode_ext"])
    sbl = frames[0]["sblimit"]
    nch = frames[0]["nch"]
    assert np.array_equal(np.stack([f["alloc"] for f in frames])[:, :nch, :sbl], g["bit_alloc"][:, :nch, :sbl])


def test_checker_rejects_corruption():
    g = np.load(os.path.join(GOLD, "Bj_S1.npz"))
    ok = g["bytes"].copy()
    for at in (0, 5, 7, 576 - 5, 576 + 2):
        bad = ok.copy()
        bad[at] ^= 0x10
        with pytest.raises(AssertionError):
            mp2_check.check_stream(bad)
