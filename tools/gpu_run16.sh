#!/bin/bash
# round-2 GPU call 16: suite + bench on the new defaults (direct release.cta hand-over, bulk ring, ring depth 3)
mkdir -p gpurun_out
( cd tools/build; echo "=== probe_pp_r0 (64 tiles)"; timeout 120 ./probe_pp_r0 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|rror" | tail -3; echo "=== probe_pp_r0 (74 tiles)"; timeout 120 ./probe_pp_r0 333 3 74 0 2>&1 | grep -E "FAIL|timing|rror" | tail -1 ) > gpurun_out/r02_probe16.log 2>&1
cat gpurun_out/r02_probe16.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest16.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest16.log; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02_pytest16.log | tail -12
timeout 300 python bench.py --kernels > gpurun_out/r02_bench16_default.json 2> gpurun_out/r02_bench16_default.err
timeout 300 python bench.py --kernels --no-cpu-baseline --batch-per-gpu 8192 > gpurun_out/r02_bench16_b8192.json 2> gpurun_out/r02_bench16_b8192.err
timeout 300 python bench.py --kernels --no-cpu-baseline --env seqslate --algo a2c --batch-per-gpu 16384 > gpurun_out/r02_bench16_c3.json 2> gpurun_out/r02_bench16_c3.err
for f in gpurun_out/r02_bench16_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
for k in d.get('kernels',[])[:9]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -11; done
