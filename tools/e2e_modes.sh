#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for opts in "panel_max_top=50000" "panel=0" "panel_max_top=30000"; do
  echo "##### GAMBA_OPTIONS=$opts"
  GAMBA_OPTIONS=$opts E2E_TIMEOUT=300 bash tools/e2e_big.sh "$@" 2>&1 | grep -E "^===|linalg \(wall\)|/tmp/sys"
done
