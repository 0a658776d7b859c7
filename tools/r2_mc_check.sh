#!/bin/bash
# Monte-Carlo combinations of the rows to reduce: parity tests, then the default bench (verbose per-matrix log)
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "monte_carlo or block_sparse or larger_systems" > gpurun_out/r2_mc_pytest.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -15 gpurun_out/r2_mc_pytest.log
t0=$SECONDS
GAMBA_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 2 --warmup 2 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/r2_mc_bench.json 2> gpurun_out/r2_mc_bench.err; echo "bench rc=$? ($((SECONDS-t0)) s)"
tail -c 2500 gpurun_out/r2_mc_bench.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r2_mc_bench.json') if l.startswith('{')][-1])
    print('bench: value %.4e e2e %.4e ms %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'],
          'frac %.3f' % j['roofline']['frac'], j['gpu_launches'], j['parity_check'], j.get('e2e_seconds'), j['engine'])
except Exception as e:
    print('no bench line', e)
PY
