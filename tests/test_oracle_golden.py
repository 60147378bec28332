"""CPU: the oracle (oracle/mp2_oracle.c) against the committed golden vectors generated from the reference
(tools/make_golden.py) -- output bytes and every decision tap, bit-exact."""
import os

import numpy as np
import pytest

import cases
import oracle

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


ALL_GOLDEN = [(1,) + g for g in cases.GOLDEN] + [(2,) + g for g in cases.GOLDEN_PSY2] + [(0,) + g for g in cases.GOLDEN_PSY0]


@pytest.mark.parametrize("psy,cfg,sig,n", ALL_GOLDEN, ids=["psy%d-%s-%s" % (p, c, s) for p, c, s, _ in ALL_GOLDEN])
def test_oracle_matches_golden(psy, cfg, sig, n):
    g = np.load(os.path.join(GOLD, ("%s_%s.npz" if psy == 1 else "psy%d_%%s_%%s.npz" % psy) % (cfg, sig)))
    fs, mode, br, pcm, pad_len, xpad = cases.make_case(cfg, sig, n)
    c = oracle.configure(fs, mode, br, psy, pad_len)
    out, tap = oracle.encode(c, pcm, xpad=xpad, taps=True)
    assert out.size == n * c.lg_frame == g["bytes"].size
    assert np.array_equal(out, g["bytes"])
    nch, sbl = c.nch, c.sblimit
    assert np.array_equal(tap["scalar"][:, :nch, :, :sbl], g["scalar"][:, :nch, :, :sbl])
    assert np.array_equal(tap["scfsi"][:, :nch, :sbl], g["scfsi"][:, :nch, :sbl])
    assert np.array_equal(tap["bit_alloc"][:, :nch, :sbl], g["bit_alloc"][:, :nch, :sbl])
    assert np.array_equal(tap["mode_ext"], g["mode_ext"]) and np.array_equal(tap["jsbound"], g["jsbound"])
    if mode == "j":
        assert np.array_equal(tap["j_scale"][:, :, :sbl], g["j_scale"][:, :, :sbl])
    assert np.array_equal(tap["smr"][:, :nch, :sbl], g["smr"][:, :nch, :sbl])  # same libm, same order: bit-exact
    # first frame: subband samples (doubles) and quantised samples
    sb = tap["sb_sample"][0].reshape(2, 3, 12, 32)
    assert np.array_equal(sb[:nch], g["sb_first"][:nch])
    alloc = tap["bit_alloc"][0]
    q = tap["q"][0].reshape(2, 3, 12, 32)
    for ch in range(nch):
        for sb_i in range(sbl):
            if alloc[ch, sb_i] and not (mode == "j" and ch == 1 and sb_i >= tap["jsbound"][0]):
                assert np.array_equal(q[ch, :, :, sb_i], g["q_first"][ch, :, :, sb_i])


def test_oracle_frame_ranges_are_independent():
    """any frame range can be produced on its own (the property the batch path shards on)"""
    fs, mode, br, pcm, _, _ = cases.make_case("Bj", "S8", 24)
    c = oracle.configure(fs, mode, br)
    full, _ = oracle.encode(c, pcm)
    for f0, f1 in ((0, 5), (5, 17), (17, 24), (23, 24)):
        part, _ = oracle.encode(c, pcm, f0, f1)
        assert np.array_equal(part, full[f0 * c.lg_frame:f1 * c.lg_frame])


def test_oracle_rejects_illegal_parameters():
    with pytest.raises(ValueError):
        oracle.configure(44000, "s", 192)
    with pytest.raises(ValueError):
        oracle.configure(48000, "s", 100)
    with pytest.raises(ValueError):
        oracle.configure(48000, "x", 192)
