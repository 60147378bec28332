#!/usr/bin/env python3
"""A/B harness: run bench.py (device-resident leg only) over (TLB_EXP, chunk size[, extra bench.py flags]) triples and
print per-kernel ns per frame (serialised) and the whole step overlapped.  TLB_EXP is an integer handed to the library
through the environment; it only means something while mp2_launch_chunk has experiment switches compiled in (the
variants measured in round 1 are listed in profiles/ncu_r1_summary.md; the losers were removed from the source).
usage: exp_sweep.py EXP:CHUNK[:--flag=..] [...]   (run on the GPU box; results appended to gpurun_out/exp_sweep.jsonl)"""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
for spec in sys.argv[1:]:
    parts = spec.split(":")
    exp, chunk = parts[0], int(parts[1])
    extra = parts[2:]  # further bench.py flags, e.g. --config=E
    env = dict(os.environ, TLB_EXP=exp)
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--no-cpu-baseline", "--no-e2e", "--steps", "4", "--warmup", "3",
           "--chunk-frames", str(chunk), "--no-dropin"] + extra
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    line = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not line:
        print(spec, "FAILED", r.stderr[-400:])
        continue
    j = json.loads(line[-1])
    ks = j["roofline"]["kernels"]
    fpl = j["roofline"]["frames_per_launch"]
    per = {k: round(v["avg_ms"] * 1e6 / fpl, 2) for k, v in ks.items()}
    rec = {"exp": exp, "chunk": chunk, "extra": extra, "value": round(j["value"]), "ns_per_frame": per,
           "sum_ns": round(sum(per.values()), 2), "serial_ns": round(j["roofline"]["serialised_ms_per_step"] * 1e6 / j["config"]["frames_per_gpu"], 2),
           "overlapped_ns": round(j["ms_per_step"] * 1e6 / j["config"]["frames_per_gpu"], 2), "parity": j.get("parity_check")}
    print(json.dumps(rec), flush=True)
    with open(os.path.join(root, "gpurun_out", "exp_sweep.jsonl"), "a") as f:
        f.write(json.dumps(rec) + "\n")
