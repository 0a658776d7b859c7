#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -c 3000 gpurun_out/bench_now.json; tail -3 gpurun_out/bench_now.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --system cyclic9-31 > gpurun_out/bench_c931.json 2>> gpurun_out/bench_now.err; tail -c 1500 gpurun_out/bench_c931.json
