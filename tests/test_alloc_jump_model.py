"""CPU: k_alloc's jump start and early exit (tests/alloc_jump_model.c, a C model of the kernel's logic) against the
oracle's verbatim greedy loop (encode_new.c:1078-1187) on random inputs: realistic SMR spreads, whole-dB values (many
exact ties between entries), identical channels, every allocation table, joint-stereo bounds, reduced budgets."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.mark.parametrize("steps,seed", [(5, 1), (4, 2), (1, 3), (8, 4)])
def test_jump_start_equals_the_plain_greedy_loop(tmp_path, steps, seed):
    exe = str(tmp_path / "alloc_jump")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "odr_audioenc_b200", "csrc"),
                    "-I" + os.path.join(ROOT, "oracle"), "-o", exe, os.path.join(HERE, "alloc_jump_model.c"), "-lm"], check=True)
    out = subprocess.run([exe, "60000", str(seed), str(steps)], capture_output=True, text=True, check=True,
                         env=dict(os.environ, HI_SPAN="64")).stdout
    assert out.strip().endswith("bad 0"), out[-2000:]
