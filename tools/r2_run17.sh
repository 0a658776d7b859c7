#!/bin/bash
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
export GAMBA_BENCH_VERBOSE=1
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel or schur or block_sparse or every_step" > gpurun_out/r2_17_quick.log 2>&1; echo "quick rc=$?" >> gpurun_out/r2_17_quick.log
tail -3 gpurun_out/r2_17_quick.log
python -c "from gamba_b200 import systems; [systems.write_system(s, '/tmp/%s.txt' % s) for s in ('cyclic9-15','kat15-15')]"
for spec in cyclic9-15:2000 kat15-15:0 kat15-15:2000; do
  sys=${spec%%:*}; sp=${spec##*:}
  GAMBA_DEBUG_TIMERS=1 gamba_b200/bin/gamba -i /tmp/$sys.txt -o /tmp/$sys.$sp.out -v 1 --max-spairs $sp > gpurun_out/r2_17_e2e_${sys}_sp$sp.log 2>&1
  echo "$sys sp=$sp $(md5sum < /tmp/$sys.$sp.out | cut -c1-32) $(grep -E 'overall|linalg \(wall|^update' gpurun_out/r2_17_e2e_${sys}_sp$sp.log | tr '\n' ' ')"
done
for n in 1 $N; do
if [ $n = 1 ]; then CMD="python bench.py --gpus 1 --steps 3 --warmup 3 --no-e2e-seconds --no-cpu-baseline"; else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3"; fi
( time $CMD > gpurun_out/r2_17_bench_n$n.json 2> gpurun_out/r2_17_bench_n$n.err ) 2>&1 | tail -3
grep -E "bench rank 0" gpurun_out/r2_17_bench_n$n.err | tail -12
done
