#!/bin/bash
cd "$GRAFT_REPO_ROOT"
ARGS="--system kat14-15 --max-spairs 0 --min-nnz 100000000 --max-matrices 3 --steps 2 --warmup 1 --no-cpu-baseline"
for n in ${NLIST:-1 2}; do
  if [ $n = 1 ]; then python bench.py $ARGS > gpurun_out/big_$n.json 2> gpurun_out/big_$n.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $n $ARGS > gpurun_out/big_$n.json 2> gpurun_out/big_$n.err; fi
  python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    l=[x for x in open(f'gpurun_out/big_{n}.json') if x.startswith('{')][-1]; j=json.loads(l)
    print(n, 'value %.3e e2e %.3e ms/step %.1f' % (j['value'], j['e2e']['value'], j['ms_per_step']), j['breakdown_ms_per_step'], j['config']['workload'])
except Exception as e:
    print(n, 'failed', e); print(open(f'gpurun_out/big_{n}.err').read()[-1500:])
PY
done
