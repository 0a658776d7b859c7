#!/bin/bash
# round-end validation on one B200: smoke(), the whole GPU test suite, the default bench and the reference arm
cd "$GRAFT_REPO_ROOT"
python __graft_entry__.py smoke 2>&1 | tail -4
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
( time python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ) 2>&1 | grep real
tail -c 4000 gpurun_out/bench_default.json
( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
tail -c 300 gpurun_out/bench_ref.json
