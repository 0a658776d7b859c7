#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GAMBA_BENCH_VERBOSE=1
python -c "from gamba_b200 import systems; systems.write_system('kat15-15', '/tmp/kat15-15.txt')"
GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/kat15-15.txt -o /tmp/k15.out -v 2 --max-spairs 0 > gpurun_out/r2_8_table_kat15-15_sp0.log 2>&1
tail -4 gpurun_out/r2_8_table_kat15-15_sp0.log
( time python bench.py --steps 3 --warmup 3 > gpurun_out/r2_8_bench.json 2> gpurun_out/r2_8_bench.err ) 2>&1 | tail -3
tail -4 gpurun_out/r2_8_bench.err
bash tools/r2_prof2.sh 5
