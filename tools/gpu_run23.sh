#!/bin/bash
# round-2 GPU call 23: FFMA2 scores epilogue: parity suite (scores-dependent tests) + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest23.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest23.log; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02_pytest23.log | tail -6
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench23_default.json 2> gpurun_out/r02_bench23_default.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench23_default.json'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), 'ms', round(d['env_only']['ms_per_step'],2))
for k in d.get('kernels',[])[:6]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
"
