#!/bin/bash
# ncu --set full of the dominant kernel (limb_gemm_tc_kernel): three late launches of the earlier-tile product (RMW_LIMBS) and the
# Schur product (RMW) of one reduction of a cyclic9-15 step matrix
cd "$(dirname "$0")/.."
IDX=${1:-5}
CMD="python tools/one_step.py cyclic9-15 15000000 $IDX"
$CMD > gpurun_out/r2_p2_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/r2_p2_plain.log; exit 1; }
cat gpurun_out/r2_p2_plain.log
# 3 reductions = 2 + 50 + 50 launches of the kernel; the last 10: tiles 16-19 of the third reduction (earlier-tile product +
# inverse-tile product each), the Schur product, the Monte-Carlo product
ncu --set full --clock-control none --import-source on -k regex:limb_gemm_tc_kernel -s 92 -c 10 -o gpurun_out/r2_p2_gemm $CMD > gpurun_out/r2_p2_ncu.log 2>&1
ls -la gpurun_out/r2_p2_gemm.ncu-rep
