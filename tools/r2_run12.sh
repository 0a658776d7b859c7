#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "from gamba_b200 import systems; [systems.write_system(s, '/tmp/%s.txt' % s) for s in ('kat15-15','eco15-31')]"
for sp in 0 2000; do
GAMBA_OPTIONS=panel=2,panel_max_gb=120,schur_max_gb=150 GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/kat15-15.txt -o /tmp/k15.$sp.out -v 2 --max-spairs $sp > gpurun_out/r2_12_kat15_sp${sp}_forced_tc.log 2>&1
echo "kat15 sp=$sp forced tc-solve: $(md5sum < /tmp/k15.$sp.out | cut -c1-32) $(grep -E 'overall|linalg \(wall' gpurun_out/r2_12_kat15_sp${sp}_forced_tc.log | tr '\n' ' ')"
grep -E "^ +[0-9]+ +[0-9]+ " gpurun_out/r2_12_kat15_sp${sp}_forced_tc.log | awk '{printf "%s %s x %s | ", $1,$4,$6; for(i=11;i<=NF;i++) printf "%s ", $i; print ""}' | tail -40
done
