"""CPU, build container only: the oracle against the reference compiled unmodified into oracle/_ref
(skipped where oracle/_ref is absent).  Longer streams than the golden fixtures; output bytes must be identical."""
import numpy as np
import pytest

import cases
import oracle
import reftool

pytestmark = pytest.mark.skipif(not reftool.have_ref(), reason="oracle/_ref not built (needs /root/reference)")

CASES = [(c, s) for c in cases.CONFIGS for s in ("S1", "S8")] + [("Bj", s) for s in cases.SIGNALS] + \
        [("C", s) for s in ("S2", "S4", "S6")] + [("A", "S6"), ("T2j", "S4")]


@pytest.mark.parametrize("cfg,sig", CASES, ids=["%s-%s" % cs for cs in CASES])
def test_oracle_bytes_equal_reference(cfg, sig):
    n = 120
    fs, mode, br, pcm, _, _ = cases.make_case(cfg, sig, n)
    c = oracle.configure(fs, mode, br)
    out, _ = oracle.encode(c, pcm)
    ref = reftool.run_ref(pcm, fs, mode, br)["bytes"]
    assert ref.size == n * c.lg_frame
    bad = np.flatnonzero((out.reshape(n, -1) != ref.reshape(n, -1)).any(axis=1))
    assert bad.size == 0, "frames differing from the reference: %s" % bad[:10]


@pytest.mark.parametrize("cfg", ["Bj", "C", "T2"])
def test_oracle_xpad_equal_reference(cfg):
    n = 60
    fs, mode, br, pcm, pad_len, xpad = cases.make_case(cfg, "PAD", n)
    c = oracle.configure(fs, mode, br, 1, pad_len)
    out, _ = oracle.encode(c, pcm, xpad=xpad)
    ref = reftool.run_ref(pcm, fs, mode, br, 1, pad_len, xpad=xpad)["bytes"]
    assert np.array_equal(out, ref)


PSY2_CASES = [("E1", "S1"), ("E1", "S8"), ("E1", "S4"), ("Bs", "S2"), ("C", "S8"), ("A", "S6"), ("M48", "S5"), ("T2j", "S8"),
              ("L2", "S1"), ("D", "S7")]


@pytest.mark.parametrize("cfg,sig", PSY2_CASES, ids=["%s-%s" % cs for cs in PSY2_CASES])
def test_oracle_psy2_equals_reference(cfg, sig):
    """psychoacoustic model 2 (inter-frame r/phi state restated as a two-block halo): bytes and SMR bit-identical"""
    n = 80
    fs, mode, br, pcm, _, _ = cases.make_case(cfg, sig, n)
    c = oracle.configure(fs, mode, br, 2)
    out, tap = oracle.encode(c, pcm, taps=True)
    r = reftool.run_ref(pcm, fs, mode, br, 2, taps=True)
    assert np.array_equal(out, r["bytes"])
    assert np.array_equal(tap["smr"][:, :c.nch], r["tap"]["smr"][:, :c.nch])


@pytest.mark.parametrize("cfg,sig", [("Bj", "S8"), ("C", "S1"), ("A", "S4"), ("T2j", "S8"), ("M48", "S6"), ("H", "S2")])
def test_oracle_psy0_equals_reference(cfg, sig):
    n = 100
    fs, mode, br, pcm, _, _ = cases.make_case(cfg, sig, n)
    c = oracle.configure(fs, mode, br, 0)
    out, tap = oracle.encode(c, pcm, taps=True)
    r = reftool.run_ref(pcm, fs, mode, br, 0, taps=True)
    assert np.array_equal(out, r["bytes"])
    assert np.array_equal(tap["smr"][:, :c.nch], r["tap"]["smr"][:, :c.nch])
