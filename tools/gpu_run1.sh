#!/bin/bash
# round-2 GPU call 1: probe every pair-kernel variant, full GPU test suite, bench A/B
mkdir -p gpurun_out
cd tools/build
for v in pair_old p2_r1_t1 p2_r0_t1 p2_r1_t0 p2_r0_t0; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./augru_probe_$v 300 1 64 1 2>&1 | tail -12
  echo "=== probe $v (300 rows div 3, 74 tiles shared)"; timeout 120 ./augru_probe_$v 300 3 74 0 2>&1 | tail -4
done > ../../gpurun_out/r02_probe1.log 2>&1
cd ../..
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02_pytest1.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest1.log
timeout 300 python bench.py --kernels > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
R4_AUGRU_PAIR=1 timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench_allpair.json 2> gpurun_out/r02_bench_allpair.err
R4_AUGRU_PAIR=1 R4_AUGRU_PAIR_IMPL=4 timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench_allpair_impl4.json 2> gpurun_out/r02_bench_allpair_impl4.err
tail -3 gpurun_out/r02_probe1.log; tail -5 gpurun_out/r02_pytest1.log
