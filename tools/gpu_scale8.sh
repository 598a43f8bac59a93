#!/bin/bash
# run under gpurun --gpus 8: the config-5 scaling points N = 8, 4, 2 (p2p exchange) and N = 8 in the NCCL form
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
for N in 8 4 2; do
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N))"
  timeout 600 $TR bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/s8_bench_$N.json 2> gpurun_out/s8_bench_$N.err; echo "bench N=$N rc=$?"
  tail -1 gpurun_out/s8_bench_$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('c5 x%d p2p: value %.4e ms %.4f e2e %.4f per-GPU roof %s | n1_same_code %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], {k:d.get('n1_same_code',{}).get(k) for k in ('value','ms_per_step')}))
print('   clocks', d['clocks'], 'nvlink_bytes', d.get('nvlink_bytes_per_step_per_gpu'))"
  grep -iE "error|Traceback" gpurun_out/s8_bench_$N.err | head -5
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
B2R_SHARD_EXCHANGE=nccl timeout 600 $TR bench.py --gpus 8 --steps 40 --warmup 5 2>/dev/null | tail -1 > gpurun_out/s8_bench_8_nccl.json
python -c "
import json
d=json.loads(open('gpurun_out/s8_bench_8_nccl.json').read())
print('c5 x8 NCCL form: value %.4e ms %.4f'%(d['value'], d['ms_per_step']))"
timeout 300 $TR tools/shard_bench.py --check --n_items 200000 --n_users 50000 --emb 128 --B 512 --K 31 --steps 3 --warmup 1 --optimizer SGD 2>&1 | grep -E "check_max_abs_err" | cut -c1-200 | head -8
