#!/bin/bash
# launch list (device time per kernel launch) of one reduction of one cyclic9-15 step matrix
cd "$(dirname "$0")/.."
IDX=${1:-5}
CMD="python tools/one_step.py cyclic9-15 15000000 $IDX"
$CMD > gpurun_out/r2_p1_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_p1_plain.log; exit 1; }
cat gpurun_out/r2_p1_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_p1_launches.csv $CMD > gpurun_out/r2_p1_ncu.log 2>&1
ls -la gpurun_out/r2_p1_launches.csv
