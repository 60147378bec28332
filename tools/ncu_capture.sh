#!/bin/bash
# GPU box, one GPU: launch list of a short bench run, then one `ncu --set full` capture of every kernel of the path
# (one launch each), exported as raw CSV for tools/ncu_traffic.py.  usage: tools/ncu_capture.sh TAG [CONFIG]
# Numbers printed by bench.py under ncu are never bench values.
set -e
TAG=${1:-r2}; CFG=${2:-B}
B="python bench.py --config $CFG --hours 0.5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin"
# 0.5 h = 75 000 frames at 48 kHz = one launch of every kernel per step; skip the three warm-up steps
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ -c 72 --csv --log-file gpurun_out/launches_${TAG}_${CFG}.csv \
    python bench.py --config $CFG --hours 1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-dropin > gpurun_out/launches_${TAG}_${CFG}.bench.log 2>&1

[ "$CFG" = "E" ] && NPER=5 || NPER=6
ncu --set full --clock-control none --import-source on -k regex:k_ -s $((3 * NPER)) -c $NPER -o gpurun_out/prof_${TAG}_${CFG} -f $B > gpurun_out/prof_${TAG}_${CFG}.log 2>&1
ncu -i gpurun_out/prof_${TAG}_${CFG}.ncu-rep --page raw --csv --print-units base > gpurun_out/prof_${TAG}_${CFG}_raw.csv
python tools/ncu_traffic.py gpurun_out/prof_${TAG}_${CFG}_raw.csv 75000 $CFG gpurun_out/ncu_${TAG}_traffic.json | tee gpurun_out/prof_${TAG}_${CFG}_table.md
