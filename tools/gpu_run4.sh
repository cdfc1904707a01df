#!/bin/bash
# round-2 GPU call 4 (2 GPUs): peer-memory gradient exchange check, bench at N=2 (peer vs NCCL), new bench modes at N=1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/r02_dist_check_n2.json 2> gpurun_out/r02_dist_check_n2.err
echo "dist_check rc $?"; tail -3 gpurun_out/r02_dist_check_n2.err; cat gpurun_out/r02_dist_check_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench4_n2.json 2> gpurun_out/r02_bench4_n2.err
echo "bench n2 rc $?"
R4_NO_PEER_COMM=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench4_n2_nccl.json 2> gpurun_out/r02_bench4_n2_nccl.err
echo "bench n2 nccl rc $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 5 --warmup 3 --sgd-minibatch 256 > gpurun_out/r02_bench4_n2_mb256.json 2> gpurun_out/r02_bench4_n2_mb256.err
echo "bench n2 strict-256 rc $?"
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --kernels > gpurun_out/r02_bench4_n1.json 2> gpurun_out/r02_bench4_n1.err &
CUDA_VISIBLE_DEVICES=1 timeout 300 python bench.py --env seqslate --algo a2c --batch-per-gpu 16384 --steps 3 --no-cpu-baseline > gpurun_out/r02_bench4_c3_seq_a2c_16384.json 2> gpurun_out/r02_bench4_c3.err
wait
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --conti --batch-per-gpu 8192 --steps 3 --no-cpu-baseline > gpurun_out/r02_bench4_c4_conti_8192.json 2> gpurun_out/r02_bench4_c4.err &
CUDA_VISIBLE_DEVICES=1 timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench4_reference.json 2> gpurun_out/r02_bench4_reference.err
wait
for f in gpurun_out/r02_bench4_*.json; do echo $f; python -c "
import json,sys
d=json.load(open('$f'))
print({k:(round(d[k]) if isinstance(d[k],float) else d[k]) for k in ('value','n_gpus') if k in d}, d.get('gradient_exchange'), 'e2e', round(d['e2e']['value']), 'env_only', round(d.get('env_only',{}).get('value',0)))
" 2>&1 | tail -1; done
