#!/bin/bash
# round-2 GPU call 21: GAE kernel + device permutation: trainer tests, bench x2
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_dataset.py -m gpu -q --timeout 600 > gpurun_out/r02_pytest21.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest21.log; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02_pytest21.log | tail -6
for i in 1 2; do timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench21_default_$i.json 2> gpurun_out/r02_bench21_default_$i.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench21_default_$i.json'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), 'ms', round(d['env_only']['ms_per_step'],2), d.get('gpu_launches'))
"; done
timeout 300 python bench.py --no-cpu-baseline --env seqslate --algo a2c --batch-per-gpu 16384 > gpurun_out/r02_bench21_c3.json 2> gpurun_out/r02_bench21_c3.err; python -c "
import json
d=json.load(open('gpurun_out/r02_bench21_c3.json'))
print('c3', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']))
"
