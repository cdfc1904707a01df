#!/bin/bash
# round-2 GPU call 30 (the last one of the round): k_reduce_adam with 16 loads in flight: trainer / parity tests,
# (the call also A/B-measured a persistent k_policy_act variant, R4_ACT_NO_WS below; that variant was never committed -- the session ended
# first -- and is NOT in the tree: the env var is inert, both lines time the committed kernel)
# then the bench lines of the round on the final kernels (default + CPU arm, batch 8192, configs[2] SeqSlate A2C 16384, configs[3] conti 8192,
# the dnn simulator at 65 536 rows)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_trainer.py tests/test_gpu_parity.py tests/test_gpu_dataset.py -m gpu -x -q > gpurun_out/r02_pytest30.log 2>&1; echo "pytest rc $?"; tail -3 gpurun_out/r02_pytest30.log
timeout 300 python bench.py --kernels > gpurun_out/r02_bench30_default.json 2> gpurun_out/r02_bench30_default.err; echo "bench rc $?"
R4_ACT_NO_WS=1 timeout 300 python bench.py --no-cpu-baseline --steps 3 > gpurun_out/r02_bench30_act_ab.json 2> gpurun_out/r02_bench30_act_ab.err; echo "bench act ab rc $?"
timeout 300 python bench.py --batch-per-gpu 8192 --no-cpu-baseline --kernels > gpurun_out/r02_bench30_b8192.json 2> gpurun_out/r02_bench30_b8192.err; echo "bench 8192 rc $?"
timeout 400 python bench.py --env seqslate --algo a2c --batch-per-gpu 16384 --no-cpu-baseline --steps 3 > gpurun_out/r02_bench30_c3.json 2> gpurun_out/r02_bench30_c3.err; echo "bench c3 rc $?"
timeout 300 python bench.py --conti --batch-per-gpu 8192 --no-cpu-baseline --steps 3 > gpurun_out/r02_bench30_c4.json 2> gpurun_out/r02_bench30_c4.err; echo "bench c4 rc $?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r02_bench30_*.json')):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value', round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), round(d['env_only']['ms_per_step'],2), 'frac', round((d.get('roofline') or {}).get('frac',0),4), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
    if 'default' in f:
        for k in d.get('kernels', []): print(k)
PY
