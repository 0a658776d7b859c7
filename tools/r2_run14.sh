#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_rowgen.py tests/test_gpu_e2e.py -x -q -m gpu > gpurun_out/r2_14_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/r2_14_quick.log
tail -8 gpurun_out/r2_14_quick.log
grep -q "rc=0" gpurun_out/r2_14_quick.log || exit 1
python -c "from gamba_b200 import systems; [systems.write_system(s, '/tmp/%s.txt' % s) for s in ('cyclic9-15','kat15-15','kat14-15','eco15-31')]"
for spec in cyclic9-15:2000 kat14-15:0 kat15-15:0 kat15-15:2000 eco15-31:0 kat15-15:0; do
  sys=${spec%%:*}; sp=${spec##*:}
  GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/$sys.txt -o /tmp/$sys.$sp.out -v 2 --max-spairs $sp > gpurun_out/r2_14_e2e_${sys}_sp$sp.log 2>&1
  echo "$sys sp=$sp $(md5sum < /tmp/$sys.$sp.out | cut -c1-32) $(grep -E 'overall|linalg \(wall|^update' gpurun_out/r2_14_e2e_${sys}_sp$sp.log | tr '\n' ' ')"
  grep -E "^dbg( stage|:)" gpurun_out/r2_14_e2e_${sys}_sp$sp.log
done
