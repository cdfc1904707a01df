#!/bin/bash
# round-2 GPU call 10: full suite (widedeep, pp forced, plug-in protocol), smoke, scores profile, bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest10.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest10.log; tail -6 gpurun_out/r02_pytest10.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke10.log 2>&1; tail -2 gpurun_out/r02_smoke10.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_scores_tc -s 12 -c 1 -f -o gpurun_out/r02_scores python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_scores.log 2>&1
tail -2 gpurun_out/r02_ncu_scores.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_gemm_tc -s 40 -c 1 -f -o gpurun_out/r02_gemm_head python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_gemm.log 2>&1
tail -2 gpurun_out/r02_ncu_gemm.log
timeout 300 python bench.py --kernels > gpurun_out/r02_bench10_default.json 2> gpurun_out/r02_bench10_default.err
timeout 300 python bench.py --simulator dnn --batch-per-gpu 65536 --steps 5 --no-cpu-baseline --kernels > gpurun_out/r02_bench10_dnn.json 2> gpurun_out/r02_bench10_dnn.err
for f in gpurun_out/r02_bench10_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
for k in d.get('kernels',[])[:9]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -11; done
