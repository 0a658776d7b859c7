#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python tools/one_step.py cyclic9-15 15000000 9 > gpurun_out/pw_plain.log 2>&1; tail -1 gpurun_out/pw_plain.log
ncu --set full --clock-control none --import-source on -k regex:elim_kernel -s 12 -c 1 -f -o gpurun_out/prof4_walk python tools/one_step.py cyclic9-15 15000000 9 > gpurun_out/pw_ncu.log 2>&1
ls -la gpurun_out/prof4_walk*
