"""CPU: the labelling k_label runs per frame-channel (tests/label_model.c, a C transcription of the kernel's logic)
against the oracle's verbatim psy-1 list code on frames of the seeded signals, among them the rare frames where the
first tonal is wiped by the second and a noise masker lands on its line (tonal and noise lists merge)."""
import os
import subprocess
import tempfile

import pytest

import signals

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("sig,n", [("S1", 1300), ("S8", 3200), ("S2", 1000)])
def test_label_model_equals_oracle(sig, n):
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "label_model")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-I" + os.path.join(ROOT, "odr_audioenc_b200", "csrc"),
                        "-I" + os.path.join(ROOT, "oracle"), "-o", exe, os.path.join(HERE, "label_model.c"), "-lm"], check=True)
        pcm = os.path.join(td, "in.pcm")
        signals.make(sig, n, 2, 48000).tofile(pcm)
        out = subprocess.run([exe, pcm, str(n)], capture_output=True, text=True, check=True).stdout
    assert out.strip().endswith("bad 0"), out[-3000:]
