#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_rowgen.py tests/test_gpu_e2e.py tests/test_gpu_ref_adapter.py -x -q -m gpu > gpurun_out/r2_9_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/r2_9_quick.log
tail -15 gpurun_out/r2_9_quick.log
grep -q "rc=0" gpurun_out/r2_9_quick.log || exit 1
python -c "from gamba_b200 import systems; [systems.write_system(s, '/tmp/%s.txt' % s) for s in ('cyclic9-15','kat15-15','kat14-15','eco15-31','cyclic9-31')]"
for spec in cyclic9-15:2000 cyclic9-31:2000 kat14-15:0 eco15-31:0 kat15-15:0 kat15-15:2000; do
  sys=${spec%%:*}; sp=${spec##*:}
  GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/$sys.txt -o /tmp/$sys.$sp.out -v 1 --max-spairs $sp > gpurun_out/r2_9_e2e_${sys}_sp$sp.log 2>&1
  echo "$sys sp=$sp $(md5sum < /tmp/$sys.$sp.out | cut -c1-32) $(grep -E 'overall|linalg \(wall|^update' gpurun_out/r2_9_e2e_${sys}_sp$sp.log | tr '\n' ' ')"
done
GAMBA_HOST_ROWGEN=1 gamba_b200/bin/gamba -i /tmp/kat15-15.txt -o /tmp/k15h.out -v 1 --max-spairs 0 > gpurun_out/r2_9_e2e_kat15-15_sp0_hostrowgen.log 2>&1
echo "host rowgen: $(grep -E 'overall|^update' gpurun_out/r2_9_e2e_kat15-15_sp0_hostrowgen.log | tr '\n' ' ')"
bash tools/r2_prof2.sh 5
