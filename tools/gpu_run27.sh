#!/bin/bash
# round-2 GPU call 27: GPU test suite + default bench on: head n-tile 256 (split-K 4), PDL chain in the PPO epoch, 75 % CTA share
# for the shared sequence in k_scores_tc2; A/B line with the head at n-tile 128 and the PPO epoch in plain stream order
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest27.log 2>&1; echo "pytest rc $?"; tail -4 gpurun_out/r02_pytest27.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench27_default.json 2> gpurun_out/r02_bench27_default.err; echo "bench rc $?"
R4_HEAD_BNT=128 R4_NO_PDL=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench27_ab.json 2> gpurun_out/r02_bench27_ab.err; echo "bench ab rc $?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench27_default.json', 'gpurun_out/r02_bench27_ab.json'):
    d=json.load(open(f))
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), round(d['env_only']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4))
    for k in d.get('kernels', []): print(k)
PY
