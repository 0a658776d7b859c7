#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GAMBA_BENCH_VERBOSE=1
timeout -k 10 900 python -m pytest tests/test_gpu_rowgen.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/r2_15_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/r2_15_quick.log
tail -4 gpurun_out/r2_15_quick.log
grep -q "rc=0" gpurun_out/r2_15_quick.log || exit 1
python -c "from gamba_b200 import systems; [systems.write_system(s, '/tmp/%s.txt' % s) for s in ('cyclic9-15','kat15-15')]"
for spec in cyclic9-15:2000 kat15-15:0; do
  sys=${spec%%:*}; sp=${spec##*:}
  GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/$sys.txt -o /tmp/$sys.$sp.out -v 2 --max-spairs $sp > gpurun_out/r2_15_e2e_${sys}_sp$sp.log 2>&1
  echo "$sys sp=$sp $(md5sum < /tmp/$sys.$sp.out | cut -c1-32) $(grep -E 'overall|linalg \(wall|^update' gpurun_out/r2_15_e2e_${sys}_sp$sp.log | tr '\n' ' ')"
  grep -E "^dbg( stage|:)" gpurun_out/r2_15_e2e_${sys}_sp$sp.log
done
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_15_bench.json 2> gpurun_out/r2_15_bench.err ) 2>&1 | tail -3
tail -4 gpurun_out/r2_15_bench.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_15_bench.json',):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['value'], d['ms_per_step'], d.get('e2e'), d.get('parity_check'))
        print(d['breakdown_ms_per_step']); print(d['roofline']['achieved'], d['roofline']['peak'], d['roofline']['frac']); print(d['engine'])
        print(d.get('e2e_seconds')); print(d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
