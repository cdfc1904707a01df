#!/bin/bash
# round-2 GPU call 5: ping-pong AUGRU kernel, dnn simulator, gathered head GEMM, new bench modes
mkdir -p gpurun_out
cd tools/build
for v in pp_r0 pp_r1; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./augru_probe_$v 300 1 64 1 2>&1 | tail -7
  echo "=== probe $v (300 rows div 3, 74 tiles shared)"; timeout 120 ./augru_probe_$v 300 3 74 0 2>&1 | tail -4
  echo "=== probe $v (300 rows div 9, 148 tiles unshared = 2 waves)"; timeout 120 ./augru_probe_$v 300 9 148 1 2>&1 | tail -2
done > ../../gpurun_out/r02_probe5.log 2>&1
cd ../..
grep -E "===|PASS|FAIL|timing|second" gpurun_out/r02_probe5.log
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_pytest5.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest5.log; tail -4 gpurun_out/r02_pytest5.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench5_default.json 2> gpurun_out/r02_bench5_default.err
R4_AUGRU_PP=1 timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench5_allpp.json 2> gpurun_out/r02_bench5_allpp.err
timeout 300 python bench.py --batch-per-gpu 8192 --kernels --no-cpu-baseline > gpurun_out/r02_bench5_b8192.json 2> gpurun_out/r02_bench5_b8192.err
timeout 300 python bench.py --env seqslate --algo a2c --batch-per-gpu 16384 --steps 3 --no-cpu-baseline > gpurun_out/r02_bench5_c3_seq_a2c_16384.json 2> gpurun_out/r02_bench5_c3.err
timeout 300 python bench.py --conti --batch-per-gpu 8192 --steps 3 --no-cpu-baseline > gpurun_out/r02_bench5_c4_conti_8192.json 2> gpurun_out/r02_bench5_c4.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench5_reference.json 2> gpurun_out/r02_bench5_reference.err
for f in gpurun_out/r02_bench5_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'frac', d.get('roofline',{}) and round(d['roofline']['frac'],3))
" 2>&1 | tail -1; done
