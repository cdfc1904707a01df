#!/bin/bash
# round-2 GPU call 29: register-blocked k_cat_attn + k_query with 16 loads in flight: GPU tests, default bench (+ side stream off for
# the A/B), ncu launch list of the bench command, ncu --set full of one k_scores_tc2 observation-pass launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest29.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest29.log
timeout 300 python bench.py --kernels --no-cpu-baseline > gpurun_out/r02_bench29_default.json 2> gpurun_out/r02_bench29_default.err; echo "bench rc $?"
R4_NO_SIDE_STREAM=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench29_noside.json 2> gpurun_out/r02_bench29_noside.err; echo "bench noside rc $?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench29_default.json', 'gpurun_out/r02_bench29_noside.json'):
    d=json.load(open(f))
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), round(d['env_only']['ms_per_step'],2), 'frac', round(d['roofline']['frac'],4))
    if 'default' in f:
        for k in d.get('kernels', []): print(k)
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_launches.log 2>&1; echo "ncu launches rc $?"
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_scores_tc2 -s 3 -c 1 -f -o gpurun_out/r02_scores_tc2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_ncu_scores_tc2.log 2>&1; echo "ncu scores rc $?"
ls -la gpurun_out/r02_scores_tc2.ncu-rep gpurun_out/r02_launches_final.csv
