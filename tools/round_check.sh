#!/bin/bash
# full GPU test suite, then the product driver end to end on named systems (md5 of the output vs recorded runs)
cd "$GRAFT_REPO_ROOT"
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
E2E_TIMEOUT=300 bash tools/e2e_big.sh "$@" > gpurun_out/e2e_round.log 2>&1
grep -E "^===|linalg \(wall\)|overall|real|/tmp/sys|update .* select|error|Error" gpurun_out/e2e_round.log
