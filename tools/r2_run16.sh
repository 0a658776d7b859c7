#!/bin/bash
# N GPUs: bench.py under torchrun (the driver's launch line), then the product driver with --gpus N on kat15-15 --max-spairs 0
cd "$(dirname "$0")/.."
N=${1:-8}
mkdir -p gpurun_out
export GAMBA_BENCH_VERBOSE=1
nvidia-smi -L | wc -l
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_16_bench_n$N.json 2> gpurun_out/r2_16_bench_n$N.err ) 2>&1 | tail -3
grep -E "bench rank 0|Error|error" gpurun_out/r2_16_bench_n$N.err | tail -8
python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/r2_16_bench_n$N.json') if l.startswith('{')][-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['parity_check'], d['breakdown_ms_per_step'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
PY
python -c "from gamba_b200 import systems; systems.write_system('kat15-15', '/tmp/kat15-15.txt')"
( time gamba_b200/bin/gamba -i /tmp/kat15-15.txt -o /tmp/k15.g$N.out -v 1 --gpus $N --max-spairs 0 ) > gpurun_out/r2_16_e2e_kat15_g$N.log 2>&1
echo "kat15-15 gpus=$N $(md5sum < /tmp/k15.g$N.out | cut -c1-32) $(grep -E 'overall|linalg \(wall' gpurun_out/r2_16_e2e_kat15_g$N.log | tr '\n' ' ')"
