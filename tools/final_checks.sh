#!/bin/bash
# round-end validation on one B200: smoke(), the whole GPU test suite, the default bench, the reference arm, e2e md5s
cd "$GRAFT_REPO_ROOT"
python __graft_entry__.py smoke 2>&1 | tail -4
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('default: value %.4e e2e %.4e ms %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'], 'frac %.3f' % j['roofline']['frac'], 'cpu %.3e' % j['cpu_baseline']['value'], j['gpu_launches'])
PY
E2E_TIMEOUT=300 bash tools/e2e_big.sh cyclic9-15 cyclic9-31 kat12-15 kat13-15 eco13-15 eco14-15 kat14-15 2>&1 | grep -E "^===|linalg \(wall\)|overall|/tmp/sys"
