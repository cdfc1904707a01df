#!/bin/bash
# round-2 GPU call 26: scores probe (grid split sweep + per-role cycles), then the GPU test suite and the default bench line on
# the new defaults (k_scores_tc2, quad-layout Kp, pay-step observation from the reward pass)
mkdir -p gpurun_out
( cd tools/build; timeout 120 ./scores_probe 4096 301 20 ) > gpurun_out/r02_probe26.log 2>&1
rc=$?; echo "probe rc $rc"; cat gpurun_out/r02_probe26.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest26.log 2>&1; echo "pytest rc $?"; tail -5 gpurun_out/r02_pytest26.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench26_default.json 2> gpurun_out/r02_bench26_default.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench26_default.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), round(d['env_only']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4))
for k in d.get('kernels', []): print(k)
PY
