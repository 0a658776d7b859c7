#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p /tmp/sys
python - <<'PY'
import sys; sys.path.insert(0,'.')
from gamba_b200 import systems
systems.write_system('kat15-15', '/tmp/sys/kat15-15.txt')
PY
( time GAMBA_DEBUG_TIMERS=1 timeout 600 gamba_b200/bin/gamba -i /tmp/sys/kat15-15.txt -o /tmp/sys/k15.out0 -v 1 --max-spairs 0 ) 2>&1 | grep -E "linalg \(wall\)|overall|update .* select|dbg:|real|rror"
md5sum /tmp/sys/k15.out0
