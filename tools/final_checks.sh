#!/bin/bash
cd "$GRAFT_REPO_ROOT"
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/bench_default.json') if l.startswith('{')][-1])
print('default: value %.3e e2e %.3e ms %.1f steps %d warmup %d' % (j['value'], j['e2e']['value'], j['ms_per_step'], j['steps'], j['warmup']), j['cpu_baseline']['value'], j['roofline']['frac'], j['gpu_launches'])
PY
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
tail -c 900 gpurun_out/bench_ref.json
( time python bench.py --system synthetic:50000x80000:0.005:31 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_syn31.json 2> gpurun_out/bench_syn31.err ) 2>&1 | grep real
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/bench_syn31.json') if l.startswith('{')][-1])
    print('syn31: value %.3e e2e %.3e ms %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'], j['roofline']['parts'])
except Exception as e:
    print('syn31 failed', e); print(open('gpurun_out/bench_syn31.err').read()[-1500:])
PY
