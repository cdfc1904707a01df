#!/bin/bash
# round-2 GPU call 11: quarter-chasing pair kernel (probe: numerics + timing + phase clocks), scores/gemm bank-conflict fix
mkdir -p gpurun_out
cd tools/build
for v in q_r1_t1 q_r0_t0 q_r0_t1 q_r1_t0 pp_r1; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./probe_$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|step 11|thread 0|rror|second" | tail -6
  echo "=== probe $v (333 rows div 3, 74 tiles shared; padding pair)"; timeout 120 ./probe_$v 333 3 74 0 2>&1 | grep -E "PASS|FAIL|timing|rror" | tail -3
done > ../../gpurun_out/r02_probe11.log 2>&1
cd ../..
cat gpurun_out/r02_probe11.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > gpurun_out/r02_pytest11.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest11.log; tail -4 gpurun_out/r02_pytest11.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench11_default.json 2> gpurun_out/r02_bench11_default.err
timeout 300 python bench.py --kernels --no-cpu-baseline --batch-per-gpu 8192 > gpurun_out/r02_bench11_b8192.json 2> gpurun_out/r02_bench11_b8192.err
for f in gpurun_out/r02_bench11_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
for k in d.get('kernels',[])[:9]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -11; done
