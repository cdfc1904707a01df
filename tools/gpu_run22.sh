#!/bin/bash
# round-2 GPU call 22: hand-over variants of the pair kernel: 0 direct release.cta quarters, 1 relayed quarters, 2 direct release.cluster whole phase, 3 relayed whole phase
mkdir -p gpurun_out
( cd tools/build; for v in probe_h_r0 probe_h_r1 probe_h_r2 probe_h_r3; do echo "=== $v (64 tiles unshared)"; timeout 120 ./$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|rror" | tail -3; echo "=== $v (74 tiles shared)"; timeout 120 ./$v 333 3 74 0 2>&1 | grep -E "FAIL|timing|rror" | tail -1; done ) > gpurun_out/r02_probe22.log 2>&1
cat gpurun_out/r02_probe22.log
