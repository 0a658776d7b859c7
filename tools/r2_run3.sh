#!/bin/bash
# quick parity (panel tests), launch list of one step, then the named-system tests with timings
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel or schur or every_step" > gpurun_out/r2_3_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/r2_3_quick.log
tail -3 gpurun_out/r2_3_quick.log
grep -q "rc=0" gpurun_out/r2_3_quick.log || exit 1
bash tools/r2_prof1.sh 5
timeout -k 10 2400 python -m pytest tests/test_gpu_named.py -q -m gpu --durations=0 > gpurun_out/r2_3_named.log 2>&1; echo "named rc=$?" >> gpurun_out/r2_3_named.log
tail -40 gpurun_out/r2_3_named.log
