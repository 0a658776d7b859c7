#!/bin/bash
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k regex:"elim_kernel" -s 700 -c 6 -f -o gpurun_out/prof_walk $CMD > gpurun_out/prof_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"schur_tc" -s 700 -c 6 -f -o gpurun_out/prof_tc $CMD > gpurun_out/prof_ncu3.log 2>&1
ls -la gpurun_out/prof_walk* gpurun_out/prof_tc*
