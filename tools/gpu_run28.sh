#!/bin/bash
# round-2 GPU call 28: k_scores_tc2 (second attention layer on the tensor pipe, A operand in TMEM) against k_scores_tc and f64
mkdir -p gpurun_out
( cd tools/build; timeout 120 ./scores_probe 4096 301 20 ) > gpurun_out/r02_probe28.log 2>&1
echo "probe rc $?"; cat gpurun_out/r02_probe28.log
