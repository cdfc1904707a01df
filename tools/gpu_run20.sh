#!/bin/bash
# round-2 GPU call 20: suite after deleting k_augru_tc; probes of the shipped kernels; sanitizer on one Slate + one lstm episode; bench
mkdir -p gpurun_out
( cd tools/build; for v in probe_pair_default probe_pp_default; do echo "=== $v (64 tiles unshared)"; timeout 120 ./$v 300 1 64 1 2>&1 | grep -E "PASS|FAIL|timing|rror|second" | tail -3; echo "=== $v (74 tiles shared)"; timeout 120 ./$v 333 3 74 0 2>&1 | grep -E "FAIL|timing|rror" | tail -1; done ) > gpurun_out/r02_probe20.log 2>&1
cat gpurun_out/r02_probe20.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02_pytest20.log 2>&1
echo "pytest rc $?" >> gpurun_out/r02_pytest20.log; grep -E "^FAILED|^ERROR|passed|failed" gpurun_out/r02_pytest20.log | tail -12
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_episode.py > gpurun_out/r02_sanitizer_memcheck_final.log 2>&1; echo "memcheck rc $?"; tail -3 gpurun_out/r02_sanitizer_memcheck_final.log
timeout 300 python bench.py --kernels > gpurun_out/r02_bench20_default.json 2> gpurun_out/r02_bench20_default.err
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench20_reference.json 2> gpurun_out/r02_bench20_reference.err
timeout 300 python bench.py --kernels --no-cpu-baseline --conti --batch-per-gpu 8192 > gpurun_out/r02_bench20_c4.json 2> gpurun_out/r02_bench20_c4.err
timeout 300 python bench.py --kernels --no-cpu-baseline --batch-per-gpu 8192 > gpurun_out/r02_bench20_b8192.json 2> gpurun_out/r02_bench20_b8192.err
for f in gpurun_out/r02_bench20_default.json gpurun_out/r02_bench20_c4.json gpurun_out/r02_bench20_b8192.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)), 'ms', round(d['env_only']['ms_per_step'],2), 'roofline', d.get('roofline',{}) and (d['roofline']['bound'], round(d['roofline']['frac'],3)))
" 2>&1 | tail -2; done; cat gpurun_out/r02_bench20_reference.json | cut -c1-400
