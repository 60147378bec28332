"""bench.py prints one JSON line with the keys the driver reads: the reference arm on CPU here, our arm on the GPU."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample-seconds", "2"], 300)
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["unit"] == "audio-s/s"
    assert "same PCM" in d["cpu_baseline"]["sample"]


def test_reference_arm_windows_come_from_the_gpu_arms_stream():
    """any window of the synthetic stream is reproducible from its absolute sample indices alone"""
    import torch
    sys.path.insert(0, ROOT)
    import bench
    cpu = torch.device("cpu")
    whole = bench.synth_pcm(0, 50000, 2, 48000, bench.STREAM_SEED, cpu)
    part = bench.synth_pcm(12345, 23456, 2, 48000, bench.STREAM_SEED, cpu)
    assert torch.equal(whole[12345:23456], part)
    assert not torch.equal(part, bench.synth_pcm(12345, 23456, 2, 48000, bench.STREAM_SEED + 1, cpu))
    cfg = bench.Cfg("B")
    wins = bench.reference_windows(cfg, 1500000, 16, 120.0)
    assert len(wins) == 16 and wins[0] == (0, 5000) and wins[-1][0] + wins[-1][1] <= 1500000
    assert cfg.flops_per_frame == 212132 + 2304 and cfg.bytes_per_frame == 5184


@pytest.mark.gpu
def test_b200_arm_line():
    d = _run(["--hours", "0.1", "--steps", "2", "--warmup", "3", "--cpu-sample-seconds", "5"], 600)
    assert BASE_KEYS | {"roofline", "clocks", "parity_check"} <= set(d)
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic", "kernel", "kernels"} <= set(r)
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["bound"] == "fp64" and r["unit"] == "TFLOP/s"
    assert abs(r["hbm"]["frac"] - r["hbm"]["achieved"] / r["hbm"]["peak"]) < 1e-9 and 0 < r["path"]["frac"] < 1
    assert d["e2e"]["copy_ceiling"]["value"] > 0 and d["dropin"]["us_per_frame"] > 0
    assert abs(sum(k["share"] for k in r["kernels"].values()) - 1.0) < 1e-6
    assert d["parity_check"]["byte_identical_to_oracle"] == d["parity_check"]["frames"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
