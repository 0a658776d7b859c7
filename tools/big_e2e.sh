#!/bin/bash
cd "$GRAFT_REPO_ROOT"
export GAMBA_DEBUG_TIMERS=1
E2E_TIMEOUT=400 bash tools/e2e_big.sh "$@" > gpurun_out/e2e_big2.log 2>&1
grep -E "^===|linalg \(wall\)|overall|real|/tmp/sys|update .* select|dbg|error|Error|reduce basis" gpurun_out/e2e_big2.log
