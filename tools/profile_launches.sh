#!/bin/bash
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
tail -c 700 gpurun_out/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/prof_launches.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
ls -la gpurun_out/prof_launches.csv
