#!/bin/bash
# ncu --set full of the elimination kernel on the largest kat14-15 steps (X = C A^-1 is ~60 % dense there)
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --system kat14-15 --min-nnz 120000000 --steps 1 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/profk_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/profk_plain.log; exit 1; }
tail -c 900 gpurun_out/profk_plain.log
ncu --set full --clock-control none --import-source on -k regex:elim_kernel -s 4 -c 2 -f -o gpurun_out/profk_elim $CMD > gpurun_out/profk_ncu.log 2>&1
ls -la gpurun_out/profk_*
