#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for c in 0 4 6 8 12 16; do
  echo "== cyclic9-15 ctas_per_sm=$c"
  timeout 300 python tools/step_stats.py cyclic9-15 15000000 ctas_per_sm=$c 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}' 
done > gpurun_out/tune_c9.log 2>&1
export STEP_STATS_MAX=4
for c in 4 8 12; do
  echo "== kat14-15 ctas_per_sm=$c"
  timeout 600 python tools/step_stats.py kat14-15 100000000 ctas_per_sm=$c 2>&1 | awk '{print $1,$2,$3,$4,$13,$14,$15,$16,$17,$18,$19}'
done > gpurun_out/tune_k14.log 2>&1
tail -80 gpurun_out/tune_c9.log; cat gpurun_out/tune_k14.log
