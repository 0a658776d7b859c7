#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_ref_adapter.py -x -q -m gpu > gpurun_out/r2_4_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/r2_4_quick.log
tail -3 gpurun_out/r2_4_quick.log
grep -q "rc=0" gpurun_out/r2_4_quick.log || exit 1
timeout -k 10 2400 python -m pytest tests/test_gpu_named.py -q -m gpu --durations=0 -k "config4 or md5" > gpurun_out/r2_4_named.log 2>&1; echo "named rc=$?" >> gpurun_out/r2_4_named.log
grep -v "^$" gpurun_out/r2_4_named.log | tail -25
# end to end phase split of the product driver on the two systems of the metric
for sys in cyclic9-15 kat15-15; do
  python -c "from gamba_b200 import systems; systems.write_system('$sys', '/tmp/$sys.txt')"
  for sp in 2000 0; do
    [ $sys = cyclic9-15 ] && [ $sp = 0 ] && continue
    GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/$sys.txt -o /tmp/$sys.out -v 1 --max-spairs $sp > gpurun_out/r2_4_e2e_${sys}_sp$sp.log 2>&1
    tail -12 gpurun_out/r2_4_e2e_${sys}_sp$sp.log
  done
done
