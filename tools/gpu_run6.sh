#!/bin/bash
# round-2 GPU call 6: FMA-pipe reciprocal variants (pair2 + ping-pong), k_policy_grad source profile, full GPU suite, dnn bench
mkdir -p gpurun_out
cd tools/build
for v in p2s0_r0_t0 p2s2_r0_t0 p2s1_r0_t0 pps0_r0 pps2_r0 pps1_r0; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./augru_probe_$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|second|timing|step 11|thread 0" | tail -8
done > ../../gpurun_out/r02_probe6.log 2>&1
cd ../..
grep -E "===|PASS|FAIL|timing|second" gpurun_out/r02_probe6.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest6.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest6.log; tail -6 gpurun_out/r02_pytest6.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_policy_grad -s 20 -c 1 -f -o gpurun_out/r02_policy_grad python tools/ppo_time.py > gpurun_out/r02_ncu_policy_grad.log 2>&1
tail -4 gpurun_out/r02_ncu_policy_grad.log
timeout 300 python bench.py --simulator dnn --batch-per-gpu 4096 --kernels --no-cpu-baseline > gpurun_out/r02_bench6_dnn_4096.json 2> gpurun_out/r02_bench6_dnn_4096.err
timeout 300 python bench.py --simulator dnn --batch-per-gpu 65536 --steps 5 --kernels > gpurun_out/r02_bench6_dnn_65536.json 2> gpurun_out/r02_bench6_dnn_65536.err
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench6_default.json 2> gpurun_out/r02_bench6_default.err
for f in gpurun_out/r02_bench6_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3), round(d['roofline']['achieved'],1)), 'hbm8d', round(d['hbm_8d']['frac_env_only'],4))
for k in d.get('kernels',[])[:8]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -10; done
