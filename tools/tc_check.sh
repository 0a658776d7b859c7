#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for o in "panel=1"; do echo "== cyclic9-31 $o"; timeout 300 python tools/step_stats.py cyclic9-31 15000000 $o 2>&1 | grep -o 'step_[0-9]*.bin\|T=[0-9]*\|elim [0-9.]* ms\|ech [0-9.]* ms' | paste - - - - | awk '{e+=$4; c+=$7; printf "%s %s %s | ", $1,$2,$4} END {print "\nSUM elim", e, "ech", c}'; done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1
E2E_TIMEOUT=300 bash tools/e2e_big.sh cyclic9-31 kat12-15 eco13-15 2>&1 | grep -E "^===|linalg \(wall\)|/tmp/sys"
