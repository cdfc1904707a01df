#!/bin/bash
# round-2 GPU call 24 (8 GPUs): exchange check + the north-star bench point (8 x 8192 = 65 536 rows) on the final kernels; N=4, N=2 points
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/dist_check.py > gpurun_out/r02_dist_check_n8_final.json 2> gpurun_out/r02_dist_check_n8_final.err
echo "dist_check rc $?"; tail -2 gpurun_out/r02_dist_check_n8_final.err; cat gpurun_out/r02_dist_check_n8_final.json | cut -c1-600
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench24_n8.json 2> gpurun_out/r02_bench24_n8.err
echo "bench n8 rc $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r02_bench24_n4.json 2> gpurun_out/r02_bench24_n4.err
echo "bench n4 rc $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench24_n2.json 2> gpurun_out/r02_bench24_n2.err
echo "bench n2 rc $?"
timeout 300 python bench.py --gpus 1 --steps 5 --warmup 3 --batch-per-gpu 8192 --no-cpu-baseline > gpurun_out/r02_bench24_n1_b8192.json 2> gpurun_out/r02_bench24_n1_b8192.err
for f in gpurun_out/r02_bench24_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print(round(d['value']), 'ms', round(d['ms_per_step'],2), d.get('gradient_exchange'), 'e2e', round(d['e2e']['value']), 'env_only', round(d['env_only']['value']), round(d['env_only']['ms_per_step'],2), 'batch', d['config']['global_batch'])
" 2>&1 | tail -1; done
