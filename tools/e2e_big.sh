#!/bin/bash
# end-to-end product driver on the larger named systems (round table + phase breakdown)
cd "$GRAFT_REPO_ROOT"
mkdir -p /tmp/sys gpurun_out
nproc; grep -m1 "model name" /proc/cpuinfo; free -g | head -2
for n in "$@"; do
python - "$n" <<'PY'
import sys; sys.path.insert(0,'.')
from gamba_b200 import systems
systems.write_system(sys.argv[1], f'/tmp/sys/{sys.argv[1]}.txt')
PY
  echo "=== $n" ; ( time timeout ${E2E_TIMEOUT:-600} gamba_b200/bin/gamba -i /tmp/sys/$n.txt -o /tmp/sys/$n.out -v 2 ) 2>&1 | tail -60
  md5sum /tmp/sys/$n.out
done
