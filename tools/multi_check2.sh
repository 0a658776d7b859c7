#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -4
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_2.json 2> gpurun_out/scale_2.err
python - <<'PY'
import json
try:
    j=json.loads([x for x in open('gpurun_out/scale_2.json') if x.startswith('{')][-1])
    print(2, 'value %.3e e2e %.3e ms/step %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'])
except Exception as e:
    print('failed', e); print(open('gpurun_out/scale_2.err').read()[-1500:])
PY
