#!/bin/bash
# --set full of the tensor-core product kernel: third reduction of the largest cyclic9-15 step (kernels chosen from history)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CMD2="python tools/one_step2.py cyclic9-15 2000 1 -1 -1"
$CMD2 > gpurun_out/r2_prof_tc_plain.log 2>&1 || { echo "plain failed"; exit 1; }
tail -1 gpurun_out/r2_prof_tc_plain.log
timeout 400 ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:"limb_gemm_tc_kernel" -s 60 -c 16 -f -o gpurun_out/r2_prof_tc $CMD2 > gpurun_out/r2_prof_tc_ncu.log 2>&1
echo "ncu tc rc=$?"; ls -la gpurun_out/r2_prof_tc.ncu-rep
ONE_STEP_REPEATS=2 timeout 300 python tools/one_step2.py kat15-15 0 3 2 1287 2>&1 | tail -2
