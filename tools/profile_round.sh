#!/bin/bash
# Round profile on one B200 (run under gpurun): plain bench first, then the ncu launch list and one
# --set full capture of the device-time leaders, all on the same short command.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/prof_plain.log; exit 1; }
tail -c 2600 gpurun_out/prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/prof_launches.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"elim_kernel" -s 2000 -c 6 -f -o gpurun_out/prof_walk $CMD > gpurun_out/prof_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"schur_tc" -s 2000 -c 6 -f -o gpurun_out/prof_tc $CMD > gpurun_out/prof_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:echelon_mma -s 40 -c 3 -f -o gpurun_out/prof_ech $CMD > gpurun_out/prof_ncu4.log 2>&1
ls -la gpurun_out/prof_*
