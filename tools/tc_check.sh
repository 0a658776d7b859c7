#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 300 python tools/step_stats.py cyclic9-15 15000000 2>&1 | awk '{print $1,$2,$3,$4,$10,$11,$13,$14,$15,$16,$17,$18,$19,$20,$21,$22}'
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_now.json') if l.startswith('{')][-1])
print('value %.3e e2e %.3e ms %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'], j['roofline']['parts']['sparse_solve_ms_per_step'], j['roofline']['frac'])
PY
GAMBA_DEBUG_TIMERS=1 E2E_TIMEOUT=300 bash tools/e2e_big.sh cyclic9-15 cyclic9-31 kat12-15 kat13-15 eco13-15 eco14-15 2>&1 | grep -E "^===|linalg \(wall\)|overall|/tmp/sys|dbg stage|dbg engine"
