#!/bin/bash
# Round-2 profile on one B200: launch list of the (short) bench command after it exited 0 without ncu, and --set full captures of
# the tensor-core product kernel on the largest cyclic9-15 step (the engine's third reduction of it: kernels chosen from history)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-e2e-seconds --no-parity"
t0=$SECONDS
$CMD > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || { echo "plain run failed"; tail -5 gpurun_out/r2_prof_plain.err; exit 1; }
echo "plain bench ok ($((SECONDS-t0)) s)"; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r2_prof_plain.json') if l.startswith('{')][-1]); print(j['value'], j['ms_per_step'], j['breakdown_ms_per_step'], j['gpu_launches'])"
t0=$SECONDS
timeout 700 ncu --target-processes application-only --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "ncu launch list rc=$? ($((SECONDS-t0)) s)"; ls -la gpurun_out/r2_launches.csv; python tools/agg_launches.py gpurun_out/r2_launches.csv all 2>&1 | tail -32
t0=$SECONDS
CMD2="python tools/one_step2.py cyclic9-15 2000 1 -1 -1"
$CMD2 > gpurun_out/r2_prof_tc_plain.log 2>&1; cat gpurun_out/r2_prof_tc_plain.log
timeout 400 ncu --target-processes application-only --set full --clock-control none --import-source on -k regex:"limb_gemm_tc_kernel" -s 110 -c 14 -f -o gpurun_out/r2_prof_tc $CMD2 > gpurun_out/r2_prof_tc_ncu.log 2>&1
echo "ncu tc rc=$? ($((SECONDS-t0)) s)"
ls -la gpurun_out/*.ncu-rep
