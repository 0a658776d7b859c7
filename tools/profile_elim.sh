#!/bin/bash
# ncu --set full of elim_kernel (Schur mode) on one cyclic9-15 step and one kat14-15 step
cd "$GRAFT_REPO_ROOT"
export STEP_STATS_MAX=6
CMD="python tools/step_stats.py cyclic9-15 15000000"
$CMD > gpurun_out/prof2_plain.log 2>&1 || { echo plain failed; tail gpurun_out/prof2_plain.log; exit 1; }
tail -3 gpurun_out/prof2_plain.log
# launches per step: stage once, reduce twice -> elim launches 2 per step; take steps 21 (4th) and 23 (6th)
ncu --set full --clock-control none --import-source on -k regex:elim_kernel -s 7 -c 1 -f -o gpurun_out/prof2_elim_c9a $CMD > gpurun_out/prof2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:elim_kernel -s 11 -c 1 -f -o gpurun_out/prof2_elim_c9b $CMD > gpurun_out/prof2_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:schur_mma -s 11 -c 1 -f -o gpurun_out/prof2_schur_c9b $CMD > gpurun_out/prof2_ncu3.log 2>&1
ls -la gpurun_out/prof2_*
