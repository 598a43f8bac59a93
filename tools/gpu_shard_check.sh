#!/bin/bash
# usage: gpu_shard_check.sh N   -- exact-update check of the sharded step only (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR tools/shard_bench.py --check --n_items 200000 --n_users 50000 --emb 128 --B 512 --K 31 --steps 3 --warmup 1 --optimizer SGD > gpurun_out/s_check_$N.log 2>&1; echo "check rc=$?" >> gpurun_out/s_check_$N.log
grep -E "check_max_abs_err|rc=" gpurun_out/s_check_$N.log | cut -c1-260 | head -12
