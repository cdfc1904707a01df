#!/bin/bash
# round-2 GPU call 3: quad-column layout (128-bit input loads) in every recurrent kernel
mkdir -p gpurun_out
cd tools/build
for v in p2q_r0_t0 p2q_r1_t1; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./augru_probe_$v 300 1 64 1 2>&1 | tail -9
  echo "=== probe $v (300 rows div 3, 74 tiles shared)"; timeout 120 ./augru_probe_$v 300 3 74 0 2>&1 | tail -3
done > ../../gpurun_out/r02_probe3.log 2>&1
echo "=== probe single (300 rows div 1, 64 tiles unshared)" >> ../../gpurun_out/r02_probe3.log
timeout 120 ./augru_probe_single 300 1 64 1 2>&1 | tail -5 >> ../../gpurun_out/r02_probe3.log
timeout 120 ./augru_probe_single 300 3 148 1 2>&1 | tail -3 >> ../../gpurun_out/r02_probe3.log
cd ../..
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_pytest3.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest3.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench3_default.json 2> gpurun_out/r02_bench3_default.err
R4_AUGRU_PAIR=1 R4_AUGRU_PAIR_IMPL=4 timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench3_allpair_impl4.json 2> gpurun_out/r02_bench3_allpair_impl4.err
R4_AUGRU_PAIR=1 timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench3_allpair.json 2> gpurun_out/r02_bench3_allpair.err
grep -E "===|timing" gpurun_out/r02_probe3.log | tail -12; tail -3 gpurun_out/r02_pytest3.log
