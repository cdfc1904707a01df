#!/bin/bash
# round-2 GPU call 15: direct-arrive semantics of the pair hand-over; suite with split-K head, 8+8-warp scores kernel, ring depth 3; bench
mkdir -p gpurun_out
cd tools/build
for v in $(ls | grep '^probe_d_' | sort); do
  echo "=== $v (64 tiles unshared)"; timeout 120 ./$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|rror" | tail -2
  echo "=== $v (74 tiles shared, padding pair)"; timeout 120 ./$v 333 3 74 0 2>&1 | grep -E "FAIL|timing|rror" | tail -1
done > ../../gpurun_out/r02_probe15.log 2>&1
cd ../..
cat gpurun_out/r02_probe15.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest15.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest15.log; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02_pytest15.log | tail -12
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench15_default.json 2> gpurun_out/r02_bench15_default.err
timeout 300 python bench.py --kernels --no-cpu-baseline --batch-per-gpu 8192 > gpurun_out/r02_bench15_b8192.json 2> gpurun_out/r02_bench15_b8192.err
for f in gpurun_out/r02_bench15_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
for k in d.get('kernels',[])[:9]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -11; done
