#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p /tmp/sys
for n in kat14-15 eco14-15 kat15-15 eco15-31; do
python - "$n" <<'PY'
import sys; sys.path.insert(0,'.')
from gamba_b200 import systems
systems.write_system(sys.argv[1], f'/tmp/sys/{sys.argv[1]}.txt')
PY
  echo "=== $n --max-spairs 0"; ( time GAMBA_DEBUG_TIMERS=1 timeout 600 gamba_b200/bin/gamba -i /tmp/sys/$n.txt -o /tmp/sys/$n.out0 -v 1 --max-spairs 0 ) 2>&1 | grep -E "linalg \(wall\)|overall|update .* select|dbg:|steps |max matrix|real|rror"
  md5sum /tmp/sys/$n.out0
done
