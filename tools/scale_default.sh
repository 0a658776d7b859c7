#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for n in ${NLIST:-4 8}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_$n.json 2> gpurun_out/scale_$n.err
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    l=[x for x in open(f'gpurun_out/scale_{n}.json') if x.startswith('{')][-1]; j=json.loads(l)
    print(n, 'value %.3e e2e %.3e ms/step %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'])
except Exception as e:
    print(n, 'failed', e); print(open(f'gpurun_out/scale_{n}.err').read()[-1500:])
PY
done
