#!/bin/bash
# round-2 GPU call 12: L2->SM ingest probe; lstm simulator + padded-LBO scores operand through the suite; bench
mkdir -p gpurun_out
timeout 300 tools/build/l2_ingest_probe > gpurun_out/r02_l2_ingest.log 2>&1; cat gpurun_out/r02_l2_ingest.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest12.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest12.log; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02_pytest12.log | tail -12
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench12_default.json 2> gpurun_out/r02_bench12_default.err
for f in gpurun_out/r02_bench12_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
for k in d.get('kernels',[])[:9]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -11; done
