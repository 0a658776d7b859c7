#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python __graft_entry__.py smoke 2>&1 | tail -5
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
GAMBA_DEBUG_TIMERS=1 E2E_TIMEOUT=400 bash tools/e2e_big.sh kat14-15 kat15-15 eco15-31 > gpurun_out/e2e_big3.log 2>&1
grep -E "^===|linalg \(wall\)|overall|/tmp/sys|dbg|update .* select|reduce basis" gpurun_out/e2e_big3.log
