#!/bin/bash
# last checks of the round on one B200: Monte-Carlo parity cases, the uncapped kat15-15 run against the oracle driver's md5, default bench
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
t0=$SECONDS
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zz_named.py -x -q -k "monte_carlo or (config4 and kat15 and extra1) or config2" > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -4 gpurun_out/r2_final_pytest.log
t0=$SECONDS
GAMBA_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err; echo "bench rc=$? ($((SECONDS-t0)) s)"
grep -v "cyclic9-15\[" gpurun_out/r2_final_bench.err | tail -c 1500
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r2_final_bench.json') if l.startswith('{')][-1])
    print('bench: value %.4e e2e %.4e ms %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'],
          'frac %.3f' % j['roofline']['frac'], j['gpu_launches'], j['parity_check'], j.get('e2e_seconds'), j['engine'], j['cpu_baseline']['value'])
except Exception as e:
    print('no bench line', e)
PY
