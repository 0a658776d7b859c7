#!/bin/bash
# ncu --set full: the Monte-Carlo row kernels on the largest sampled kat15-15 step, the tensor-core product on the largest cyclic9-15 step
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CMD="python tools/one_step2.py kat15-15 0 3 2 1287 mc=1"
ONE_STEP_REPEATS=2 $CMD > gpurun_out/r2_prof_mc_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_prof_mc_plain.log; exit 1; }
cat gpurun_out/r2_prof_mc_plain.log
ONE_STEP_REPEATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"mc_rows_kernel|elim_block_kernel" -c 2 -f -o gpurun_out/r2_prof_mc $CMD > gpurun_out/r2_prof_mc_ncu.log 2>&1
echo "ncu mc rc=$?"
CMD2="python tools/one_step2.py cyclic9-15 2000 1 -1 -1 mc=0"
$CMD2 > gpurun_out/r2_prof_tc_plain.log 2>&1; cat gpurun_out/r2_prof_tc_plain.log
ONE_STEP_REPEATS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"limb_gemm_tc_kernel" -s 30 -c 12 -f -o gpurun_out/r2_prof_tc $CMD2 > gpurun_out/r2_prof_tc_ncu.log 2>&1
echo "ncu tc rc=$?"
ls -la gpurun_out/*.ncu-rep
