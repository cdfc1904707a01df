#!/bin/bash
# round-2 GPU call 8: cluster-shared (multicast) weight stream in the pair kernel
mkdir -p gpurun_out
cd tools/build
for v in cs2_r0 cs4_r0 cs8_r0 cs4_r1 cs8_r1; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./augru_probe_$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|step 11|thread 0|rror" | tail -6
  echo "=== probe $v (333 rows div 3, 74 tiles shared; 3 row tiles -> padding pair)"; timeout 120 ./augru_probe_$v 333 3 74 0 2>&1 | grep -E "PASS|FAIL|timing|rror" | tail -3
done > ../../gpurun_out/r02_probe8.log 2>&1
cd ../..
cat gpurun_out/r02_probe8.log
timeout 900 python -m pytest tests/test_gpu_dnn.py tests/test_gpu_parity_regimes.py -m gpu -q --timeout 900 > gpurun_out/r02_pytest8.log 2>&1
echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest8.log
for cs in 2 4 8; do
R4_AUGRU_CLUSTER=$cs timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench8_cs$cs.json 2> gpurun_out/r02_bench8_cs$cs.err
R4_AUGRU_CLUSTER=$cs R4_AUGRU_PAIR=1 timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench8_allpair_cs$cs.json 2> gpurun_out/r02_bench8_allpair_cs$cs.err
done
R4_AUGRU_CLUSTER=4 R4_AUGRU_PAIR=1 timeout 300 python bench.py --batch-per-gpu 8192 --kernels --no-cpu-baseline > gpurun_out/r02_bench8_b8192_allpair_cs4.json 2> gpurun_out/r02_bench8_b8192.err
for f in gpurun_out/r02_bench8_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
for k in d.get('kernels',[])[:4]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -5; done
