#!/bin/bash
# round-2 GPU call 7: epilogue ablation probes, GPU suite, bench after scores / head / policy_grad changes, sanitizer
mkdir -p gpurun_out
cd tools/build
for v in abl0 abl1 abl2 abl4 abl8 abl15; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./augru_probe_$v 300 1 64 1 2>&1 | grep -E "timing|step 11|thread 0" | tail -5
done > ../../gpurun_out/r02_probe7.log 2>&1
cd ../..
cat gpurun_out/r02_probe7.log
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest7.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest7.log; tail -5 gpurun_out/r02_pytest7.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench7_default.json 2> gpurun_out/r02_bench7_default.err
timeout 300 python bench.py --batch-per-gpu 8192 --kernels --no-cpu-baseline > gpurun_out/r02_bench7_b8192.json 2> gpurun_out/r02_bench7_b8192.err
for f in gpurun_out/r02_bench7_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
for k in d.get('kernels',[])[:9]: print('    %-44s %8.3f ms x%d'%(k['name'],k['ms'],k['launches']))
" 2>&1 | tail -11; done
timeout 900 compute-sanitizer --tool memcheck --log-file gpurun_out/r02_memcheck.log python tools/sanitize_episode.py 160 > gpurun_out/r02_memcheck.out 2>&1
echo "memcheck rc $?"; tail -3 gpurun_out/r02_memcheck.log; tail -3 gpurun_out/r02_memcheck.out
timeout 900 compute-sanitizer --tool racecheck --log-file gpurun_out/r02_racecheck.log python tools/sanitize_episode.py 160 > gpurun_out/r02_racecheck.out 2>&1
echo "racecheck rc $?"; tail -3 gpurun_out/r02_racecheck.log; tail -3 gpurun_out/r02_racecheck.out
