#!/bin/bash
# usage: gpu_shard.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/s_gpus_$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR tools/shard_bench.py --check --n_items 200000 --n_users 50000 --emb 128 --B 512 --K 31 --steps 5 --warmup 2 --optimizer SGD > gpurun_out/s_check_$N.log 2>&1; echo "check rc=$?" >> gpurun_out/s_check_$N.log
grep -E "check_max_abs_err|rc=|Error|error" gpurun_out/s_check_$N.log | head -12
timeout 600 $TR tools/shard_bench.py --n_items 100000000 --n_users 1000000 --emb 128 --B 4096 --K 255 --steps 30 --warmup 5 --torch_profile > gpurun_out/s_c5_$N.log 2>&1; echo "c5 rc=$?" >> gpurun_out/s_c5_$N.log
grep -E '^\{|rc=|Error|error|OutOfMemory' gpurun_out/s_c5_$N.log | cut -c1-330 | head -8
if [ "$2" != "nobench" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 300 --warmup 10 > gpurun_out/s_bench_c2_$N.json 2> gpurun_out/s_bench_c2_$N.err; echo "bench rc=$?"
tail -1 gpurun_out/s_bench_c2_$N.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('c2 replicas x$N value %.3e ms %.4f e2e %.3e'%(d['value'], d['ms_per_step'], d['e2e']['value']))"
fi
