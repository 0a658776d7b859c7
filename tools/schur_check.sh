#!/bin/bash
# parity of the tensor-core product / blocked elimination + per-step timings
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "schur or sparse_path or blocked" > gpurun_out/schur_pytest.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/schur_pytest.log
timeout 600 python tools/step_stats.py cyclic9-15 15000000 > gpurun_out/schur_c9_on.log 2>&1; tail -12 gpurun_out/schur_c9_on.log
STEP_STATS_MAX=4 timeout 600 python tools/step_stats.py kat14-15 100000000 > gpurun_out/schur_k14_on.log 2>&1; tail -5 gpurun_out/schur_k14_on.log
