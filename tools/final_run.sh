#!/bin/bash
# GPU box, one GPU: the evidence set of a round (TAG = r2): ncu launch list + full capture (-> traffic JSON that the
# bench line quotes), the default bench line and the reference arm, configs C and E, parity sweeps, probes.
TAG=${1:-r2}
bash tools/ncu_capture.sh $TAG B > gpurun_out/ncu_capture_$TAG.log 2>&1
cp gpurun_out/ncu_${TAG}_traffic.json profiles/ncu_r2_traffic.json   # bench.py reads it from profiles/
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${TAG}_final.json 2> gpurun_out/bench_${TAG}_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${TAG}_final_reference.json 2>> gpurun_out/bench_${TAG}_final.err
python bench.py --config C --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_final_cfgC.json 2>> gpurun_out/bench_${TAG}_final.err
python bench.py --config E --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_final_cfgE.json 2>> gpurun_out/bench_${TAG}_final.err
python tools/parity_sweep.py ${SWEEP_N:-20000} gpurun_out/parity_sweep_${TAG}_final.json > gpurun_out/parity_sweep_${TAG}_final.log 2>&1
SWEEP_PSY=2 python tools/parity_sweep.py ${SWEEP_N2:-10000} gpurun_out/parity_sweep_psy2_${TAG}_final.json > gpurun_out/parity_sweep_psy2_${TAG}_final.log 2>&1
SWEEP_PSY=0 python tools/parity_sweep.py 2000 gpurun_out/parity_sweep_psy0_${TAG}_final.json > gpurun_out/parity_sweep_psy0_${TAG}_final.log 2>&1
python tools/probes.py > gpurun_out/probes_${TAG}.json 2>/dev/null
tail -q -n 1 gpurun_out/parity_sweep_${TAG}_final.log gpurun_out/parity_sweep_psy2_${TAG}_final.log gpurun_out/parity_sweep_psy0_${TAG}_final.log
python - <<PY
import json
for f in ("final", "final_reference", "final_cfgC", "final_cfgE"):
    try:
        d = json.loads(open("gpurun_out/bench_${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], (d.get("roofline") or {}).get("frac"), ((d.get("roofline") or {}).get("path") or {}).get("frac"), d.get("parity_check"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -n 3 gpurun_out/bench_${TAG}_final.err
