#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "from gamba_b200 import systems; systems.write_system('kat15-15', '/tmp/kat15-15.txt'); systems.write_system('cyclic9-15', '/tmp/cyclic9-15.txt')"
GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/cyclic9-15.txt -o /tmp/c9.out -v 2 > gpurun_out/r2_5_table_cyclic9-15.log 2>&1
GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/kat15-15.txt -o /tmp/k15.out -v 2 > gpurun_out/r2_5_table_kat15-15.log 2>&1
tail -3 gpurun_out/r2_5_table_kat15-15.log
( time python bench.py --steps 3 --warmup 3 > gpurun_out/r2_5_bench.json 2> gpurun_out/r2_5_bench.err ) 2>&1 | tail -3; echo "bench rc=$?"
tail -3 gpurun_out/r2_5_bench.err
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_5_bench_ref.json 2> gpurun_out/r2_5_bench_ref.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2_5_bench_ref.err
python - <<'PY'
import json
for f in ('gpurun_out/r2_5_bench.json','gpurun_out/r2_5_bench_ref.json'):
    try:
        d=json.load(open(f)); print(f, d['value'], d['ms_per_step'], d.get('e2e'), d.get('parity_check'), d.get('breakdown_ms_per_step'))
        r=d.get('roofline');
        if r: print(r['achieved'], r['peak'], r['frac'], r['hbm_alg']['frac'])
        print(d.get('e2e_seconds')); print(d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
