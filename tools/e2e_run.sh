#!/bin/bash
# end-to-end product driver on the named systems, phase breakdown
cd "$GRAFT_REPO_ROOT"
mkdir -p /tmp/sys gpurun_out
python - <<'PY'
import sys; sys.path.insert(0,'.')
from gamba_b200 import systems
for n in ['cyclic9-15','cyclic9-31','kat12-15','eco12-15','kat13-15','eco13-15']:
    systems.write_system(n, f'/tmp/sys/{n}.txt')
PY
for n in cyclic9-15 cyclic9-31 kat12-15 eco12-15 kat13-15 eco13-15; do
  echo "=== $n" ; ( time timeout 300 gamba_b200/bin/gamba -i /tmp/sys/$n.txt -o /tmp/sys/$n.out -v 1 ) 2>&1 | tail -14
  md5sum /tmp/sys/$n.out
done
