#!/bin/bash
# round 2, first GPU run: the blocked tensor-core solve against the oracle, then the whole GPU suite, then a bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_1_gpu.txt; nproc >> gpurun_out/r2_1_gpu.txt
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel" > gpurun_out/r2_1_panel.log 2>&1; echo "panel rc=$?" >> gpurun_out/r2_1_panel.log
tail -5 gpurun_out/r2_1_panel.log
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_1_pytest.log
tail -5 gpurun_out/r2_1_pytest.log
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err; echo "bench rc=$?"
cat gpurun_out/r2_1_bench.json | head -c 3000
