#!/bin/bash
# round-2 GPU call 13: what the pair kernel's MMA warp waits for (operand quarters vs ring stages); ring depth; no-allocate loads
mkdir -p gpurun_out
cd tools/build
for v in w_d w_nst4 w_nst3 w_noal; do
  echo "=== probe $v (300 rows div 1, 64 tiles unshared)"; timeout 120 ./probe_$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|step 1[012]|thread 0|rror" | tail -9
done > ../../gpurun_out/r02_probe13.log 2>&1
cd ../..
cat gpurun_out/r02_probe13.log
