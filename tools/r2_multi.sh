#!/bin/bash
# 2 GPUs: N > 1 parity tests, then the default bench under torchrun (short)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
nvidia-smi -L
t0=$SECONDS
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2_multi_pytest.log 2>&1; echo "pytest multi rc=$? ($((SECONDS-t0)) s)"; tail -5 gpurun_out/r2_multi_pytest.log
t0=$SECONDS
N=${NGPU:-2}
GAMBA_BENCH_VERBOSE=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 2 --warmup 2 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench N=$N rc=$? ($((SECONDS-t0)) s)"
grep -v "^\[bench rank [1-9]" gpurun_out/r2_bench_n$N.err | tail -c 1800
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/r2_bench_n$N.json') if l.startswith('{')][-1])
    print('bench N=%d: value %.4e e2e %.4e ms %.1f' % (j['n_gpus'], j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'], j['parity_check'], j['engine']['solve_modes'])
except Exception as e:
    print('no bench line', e)
PY
