#!/bin/bash
# parity of the tensor-core product + per-step timings with and without it
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "schur or sparse_path" > gpurun_out/schur_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/schur_pytest.log
timeout 600 python tools/step_stats.py cyclic9-15 15000000 schur=1 > gpurun_out/schur_c9_on.log 2>&1; tail -12 gpurun_out/schur_c9_on.log
timeout 600 python tools/step_stats.py cyclic9-15 15000000 schur=0 > gpurun_out/schur_c9_off.log 2>&1; tail -12 gpurun_out/schur_c9_off.log
