#!/bin/bash
# GPU suite, then the launch list of one cyclic9-15 step and a short bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_2_pytest.log
tail -4 gpurun_out/r2_2_pytest.log
grep -q "rc=0" gpurun_out/r2_2_pytest.log || exit 1
bash tools/r2_prof1.sh 5
timeout -k 10 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_2_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['breakdown_ms_per_step'], d['roofline']['parts']['sparse_solve_ms_per_step'], d['roofline']['parts']['schur_ms_per_step'])
PY
