#!/bin/bash
# round-2 GPU call 2: epilogue variants of the pair kernel, ncu source-level capture, batch-8192 launch list
mkdir -p gpurun_out
cd tools/build
for v in p2x_ps0_sw0 p2x_ps1_sw0 p2x_ps0_sw1 p2x_ps1_sw1; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./augru_probe_$v 300 1 64 1 2>&1 | tail -9
done > ../../gpurun_out/r02_probe2.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_augru_pair2 -s 1 -c 1 -f -o ../../gpurun_out/r02_pair2_r0t0 ./augru_probe_p2x_ps0_sw0 300 1 64 1 > ../../gpurun_out/r02_ncu_pair2.log 2>&1
cd ../..
timeout 300 python bench.py --batch-per-gpu 8192 --kernels --no-cpu-baseline > gpurun_out/r02_bench_b8192.json 2> gpurun_out/r02_bench_b8192.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_b8192.csv python bench.py --batch-per-gpu 8192 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_b8192.log 2>&1
tail -3 gpurun_out/r02_probe2.log
