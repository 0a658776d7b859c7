#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/tc_pytest.log 2>&1
echo "pytest rc=$?"; tail -12 gpurun_out/tc_pytest.log
for m in 1 0; do echo "== mc_ech=$m"; timeout 300 python tools/step_stats.py cyclic9-15 15000000 mc_ech=$m 2>&1 | grep -o 'step_[0-9]*.bin\|new [0-9]*\|elim [0-9.]* ms\|ech [0-9.]* ms' | paste - - - -; done
