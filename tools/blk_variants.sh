#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "blocked or schur_every" 2>&1 | tail -2
timeout 300 python tools/step_stats.py cyclic9-15 15000000 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}'
STEP_STATS_MAX=4 timeout 300 python tools/step_stats.py kat14-15 100000000 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}'
