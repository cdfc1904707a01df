#!/bin/bash
# round-2 GPU call 18: ncu of the final pair kernel (probe) and of the 8+8-warp scores kernel (bench); sanity bench
mkdir -p gpurun_out
( cd tools/build; timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_augru_pair2 -s 1 -c 1 -f -o ../../gpurun_out/r02_pair2_final ./probe_pair_default 300 1 64 1 > ../../gpurun_out/r02_ncu_pair2_final.log 2>&1 )
tail -2 gpurun_out/r02_ncu_pair2_final.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scores_tc -s 13 -c 1 -f -o gpurun_out/r02_scores_v2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_scores_v2.log 2>&1
tail -2 gpurun_out/r02_ncu_scores_v2.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench18_default.json 2> gpurun_out/r02_bench18_default.err
python -c "
import json
d=json.load(open('gpurun_out/r02_bench18_default.json'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), 'ms', round(d['env_only']['ms_per_step'],2))
"
