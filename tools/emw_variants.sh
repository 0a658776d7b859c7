#!/bin/bash
cd "$GRAFT_REPO_ROOT"
cp gamba_b200/libgamba_cuda.so /tmp/lib_orig.so
for v in 512 256 128; do
  if [ $v = 512 ]; then cp /tmp/lib_orig.so gamba_b200/libgamba_cuda.so; else cp gamba_b200/lib_emw$v.so gamba_b200/libgamba_cuda.so; fi
  echo "== EM_W=$v"
  true
  timeout 300 python tools/step_stats.py cyclic9-15 15000000 2>&1 | grep -o 'step_[0-9]*.bin\|new [0-9]*\|ech [0-9.]* ms' | paste - - -
done
cp /tmp/lib_orig.so gamba_b200/libgamba_cuda.so
