"""CPU: the mask-driven tonal walk of k_label is equivalent to the reference's linked-list walk
(tests/tonal_walk_model.c: 200 000 random spectra incl. ties and tonals closer than `run`)."""
import os
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def test_mask_walk_equals_list_walk():
    with tempfile.TemporaryDirectory() as td:
        exe = os.path.join(td, "walk")
        subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, os.path.join(HERE, "tonal_walk_model.c"), "-lm"], check=True)
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert out.strip().endswith("bad 0"), out[-2000:]
