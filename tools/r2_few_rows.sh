#!/bin/bash
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "warps_per_row or monte_carlo" 2>&1 | tail -4
timeout 400 python tools/mc_seq.py kat15-15 0 3 3 2>&1 | grep -E "pass [12] step [1-4] " | cut -c1-200
