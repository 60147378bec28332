#!/bin/bash
# GPU box with N GPUs: the two strong-scaling workloads of SURVEY 8(e) at world size N (one JSON line each), plus the
# weak-scaling default when WEAK=1.  usage: tools/scale_run.sh N TAG
N=$1; TAG=${2:-r2}
if [ "$N" = 1 ]; then T="python"; else T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"; fi
$T bench.py --gpus $N --strong --steps 5 --warmup 3 --no-cpu-baseline --no-dropin > gpurun_out/bench_${TAG}_strongB_n$N.json 2> gpurun_out/bench_${TAG}_strongB_n$N.err
$T bench.py --gpus $N --config D --steps 5 --warmup 3 > gpurun_out/bench_${TAG}_cfgD_n$N.json 2> gpurun_out/bench_${TAG}_cfgD_n$N.err
if [ "$WEAK" = 1 ]; then
  $T bench.py --gpus $N --steps 5 --warmup 3 --no-dropin > gpurun_out/bench_${TAG}_weakB_n$N.json 2> gpurun_out/bench_${TAG}_weakB_n$N.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${TAG}_*_n$N.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "FAILED", e); print(open(f.replace(".json", ".err")).read()[-800:]); continue
    e = d.get("e2e") or {}
    print(f.split("/")[-1], "value %.0f" % d["value"], "e2e %.0f" % e.get("value", 0), "ceiling", (e.get("copy_ceiling") or {}).get("value"), d.get("parity_check"))
PY
