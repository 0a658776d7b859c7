#!/bin/bash
# 2 GPUs: N>1 parity tests, the product driver with --gpus 2, bench.py at N=2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_6_gpus.txt
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_6_multi.log 2>&1; echo "multi rc=$?" >> gpurun_out/r2_6_multi.log
tail -15 gpurun_out/r2_6_multi.log
python -c "from gamba_b200 import systems; systems.write_system('cyclic9-15', '/tmp/cyclic9-15.txt'); systems.write_system('kat14-15', '/tmp/kat14-15.txt')"
for g in 1 2; do
  for sys in cyclic9-15 kat14-15; do
    ( time gamba_b200/bin/gamba -i /tmp/$sys.txt -o /tmp/$sys.g$g.out -v 1 --gpus $g --max-spairs 0 ) > gpurun_out/r2_6_e2e_${sys}_g$g.log 2>&1
    echo "$sys gpus=$g $(md5sum < /tmp/$sys.g$g.out) $(grep -E 'overall|linalg \(wall' gpurun_out/r2_6_e2e_${sys}_g$g.log | tr '\n' ' ')"
  done
done
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_6_bench_n2.json 2> gpurun_out/r2_6_bench_n2.err ) 2>&1 | tail -3
tail -5 gpurun_out/r2_6_bench_n2.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r2_6_bench_n2.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['parity_check'], d['breakdown_ms_per_step'], d['roofline']['frac'])
except Exception as e: print('ERR', e)
PY
