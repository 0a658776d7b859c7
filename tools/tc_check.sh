#!/bin/bash
cd "$GRAFT_REPO_ROOT"
GAMBA_DEBUG_TIMERS=1 E2E_TIMEOUT=300 bash tools/e2e_big.sh cyclic9-31 > gpurun_out/e2e_dbg.log 2>&1
tail -75 gpurun_out/e2e_dbg.log | cut -c1-200
