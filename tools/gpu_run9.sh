#!/bin/bash
# round-2 GPU call 9 (8 GPUs): exchange check at N=8 and the north-star bench point (8 x 8192 = 65 536 rows)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py > gpurun_out/r02_dist_check_n8.json 2> gpurun_out/r02_dist_check_n8.err
echo "dist_check rc $?"; tail -2 gpurun_out/r02_dist_check_n8.err; cat gpurun_out/r02_dist_check_n8.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench9_n8.json 2> gpurun_out/r02_bench9_n8.err
echo "bench n8 rc $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 3 --warmup 3 --env seqslate > gpurun_out/r02_bench9_n8_seqslate.json 2> gpurun_out/r02_bench9_n8_seqslate.err
echo "bench n8 seqslate rc $?"
for f in gpurun_out/r02_bench9_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), d.get('gradient_exchange'), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), round(d['env_only']['ms_per_step'],2), 'batch', d['config']['global_batch'])
" 2>&1 | tail -1; done
