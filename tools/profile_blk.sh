#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export STEP_STATS_MAX=6
CMD="python tools/step_stats.py cyclic9-15 15000000"
$CMD > gpurun_out/prof3_plain.log 2>&1 || { echo plain failed; tail gpurun_out/prof3_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:elim_block -s 11 -c 1 -f -o gpurun_out/prof3_blk_c9 $CMD > gpurun_out/prof3_ncu1.log 2>&1
ls -la gpurun_out/prof3_*
