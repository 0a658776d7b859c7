#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "schur or blocked or sparse_path" > gpurun_out/tc_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/tc_pytest.log
for tc in 1 0; do echo "== schur_tc=$tc"
timeout 300 python tools/step_stats.py cyclic9-15 15000000 schur_tc=$tc 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}'
STEP_STATS_MAX=3 timeout 300 python tools/step_stats.py kat14-15 100000000 schur_tc=$tc 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}'
done
STEP_STATS_MAX=4 timeout 300 python tools/step_stats.py cyclic9-31 15000000 schur_tc=1 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}'
STEP_STATS_MAX=4 timeout 300 python tools/step_stats.py cyclic9-31 15000000 schur_tc=0 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}'
