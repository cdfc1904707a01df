#!/bin/bash
# round-2 GPU call 14: ring depth x no-allocate input loads x hand-over variant for the pair kernel; ping-pong kernel ring depth
mkdir -p gpurun_out
cd tools/build
for v in $(ls | grep '^probe_m_\|^probe_pp_' | sort); do
  echo "=== $v (64 tiles unshared)"; timeout 120 ./$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|rror" | tail -2
  echo "=== $v (74 tiles shared, padding pair)"; timeout 120 ./$v 333 3 74 0 2>&1 | grep -E "FAIL|timing|rror" | tail -1
done > ../../gpurun_out/r02_probe14.log 2>&1
cd ../..
cat gpurun_out/r02_probe14.log
