#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GAMBA_BENCH_VERBOSE=1
( time python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_10_bench.json 2> gpurun_out/r2_10_bench.err ) 2>&1 | tail -3
tail -3 gpurun_out/r2_10_bench.err
D=/tmp/gamba_b200_bench/kat15-15_sp0_min0_every3
ls -la $D | head -12
python tools/block_occupancy.py $D/step_0006.bin $D/step_0009.bin $D/step_0012.bin > gpurun_out/r2_10_occupancy.txt 2>&1
cat gpurun_out/r2_10_occupancy.txt
D2=/tmp/gamba_b200_bench/cyclic9-15_sp2000_min0_every1
python tools/block_occupancy.py $D2/step_0023.bin >> gpurun_out/r2_10_occupancy.txt 2>&1
tail -7 gpurun_out/r2_10_occupancy.txt
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_10_bench_ref.json 2> gpurun_out/r2_10_bench_ref.err ) 2>&1 | tail -3
python - <<'PY'
import json
for f in ('gpurun_out/r2_10_bench.json','gpurun_out/r2_10_bench_ref.json'):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1]); print(f, d['value'], d['ms_per_step'], d.get('e2e'), d.get('parity_check'))
        print(d.get('e2e_seconds')); print(d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
