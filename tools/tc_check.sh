#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for w in 4 8; do echo "== elim_warps=$w"; timeout 300 python tools/step_stats.py cyclic9-15 15000000 elim_warps=$w 2>&1 | grep -o 'step_[0-9]*.bin\|T=[0-9]*\|sparse [0-9.]*' | paste - - -; done
