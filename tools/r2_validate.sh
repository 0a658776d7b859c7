#!/bin/bash
# round-2 validation on one B200: smoke(), default bench (both arms), the GPU test suite with durations, launch list.
# Every part has its own timeout and log under gpurun_out/ so a failure in one does not lose the others.
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo; nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader
t0=$SECONDS
timeout 300 python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$? ($((SECONDS-t0)) s)"; tail -5 gpurun_out/r2_smoke.log
t0=$SECONDS
GAMBA_BENCH_VERBOSE=1 timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; echo "bench rc=$? ($((SECONDS-t0)) s)"
tail -c 1500 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r2_bench_default.json') if l.startswith('{')][-1])
    print('default: value %.4e e2e %.4e ms %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'],
          'frac %.3f' % j['roofline']['frac'], j['gpu_launches'], j['parity_check'], j.get('e2e_seconds'), j['engine'])
except Exception as e:
    print('no bench line', e)
PY
if [ "${SKIP_REF:-0}" != 1 ]; then
t0=$SECONDS
timeout 400 python bench.py --impl reference > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "reference arm rc=$? ($((SECONDS-t0)) s)"
head -c 600 gpurun_out/r2_bench_reference.json; echo
fi
if [ "${SKIP_TESTS:-0}" != 1 ]; then
t0=$SECONDS
timeout 1500 python -m pytest tests -x -q -m gpu --durations=25 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$? ($((SECONDS-t0)) s)"; tail -40 gpurun_out/r2_pytest_gpu.log
fi
if [ "${SKIP_NCU:-0}" != 1 ]; then
t0=$SECONDS
CMD="python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-e2e-seconds --no-parity"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launch list rc=$? ($((SECONDS-t0)) s)"; ls -la gpurun_out/r2_launches.csv; python tools/agg_launches.py gpurun_out/r2_launches.csv all 2>&1 | tail -30
fi
