#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export GAMBA_BENCH_VERBOSE=1
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 3 --systems cyclic9-15:2000:1 --min-nnz 15000000 > gpurun_out/r2_7_bench_n2.json 2> gpurun_out/r2_7_bench_n2.err ) 2>&1 | tail -3
echo "rc=$?"; tail -20 gpurun_out/r2_7_bench_n2.err; head -c 1500 gpurun_out/r2_7_bench_n2.json
