#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for t in 0 3072 4096 6144; do echo "== panel_t=$t"; timeout 300 python tools/step_stats.py cyclic9-15 15000000 panel_t=$t 2>&1 | grep -o 'step_[0-9]*.bin\|T=[0-9]*\|sparse [0-9.]*' | paste - - - | awk '{s+=$4; printf "%s %s %s | ", $1,$2,$4} END {print "\nSUM sparse", s}'; done
