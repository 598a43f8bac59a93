#!/bin/bash
# usage: gpu_shard2.sh N   (run under gpurun --gpus N): p2p exchange -- correctness at small sizes, then the c5 bench line
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for EX in p2p nccl; do
  B2R_SHARD_EXCHANGE=$EX timeout 300 $TR tools/shard_bench.py --check --n_items 200000 --n_users 50000 --emb 128 --B 512 --K 31 --steps 5 --warmup 2 --optimizer SGD > gpurun_out/s2_check_${EX}_$N.log 2>&1; echo "check $EX rc=$?"
  grep -E "check_max_abs_err|Error|error|Traceback" gpurun_out/s2_check_${EX}_$N.log | cut -c1-250 | head -6
done
# Adam, config-5 shapes at reduced table size: p2p and nccl forms must agree with each other (same seeds)
for EX in p2p nccl; do
  B2R_SHARD_EXCHANGE=$EX timeout 300 $TR tools/shard_bench.py --n_items 4000000 --n_users 100000 --emb 128 --B 4096 --K 255 --steps 10 --warmup 3 > gpurun_out/s2_adam_${EX}_$N.log 2>&1
  grep -E '^\{' gpurun_out/s2_adam_${EX}_$N.log | cut -c1-200
done
timeout 900 $TR bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/s2_bench_$N.json 2> gpurun_out/s2_bench_$N.err; echo "bench rc=$?"
tail -1 gpurun_out/s2_bench_$N.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('c5 x%d: value %.3e ms %.4f e2e %.4f roof %s exchange: %s'%(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['config']['exchange'][:60]))
print('n1_same_code', d.get('n1_same_code'))
print('nvlink bytes', d.get('nvlink_bytes_per_step_per_gpu'))
"
tail -5 gpurun_out/s2_bench_$N.err | cut -c1-300
B2R_SHARD_EXCHANGE=nccl timeout 600 $TR bench.py --gpus $N --steps 40 --warmup 5 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('c5 x%d NCCL form: value %.3e ms %.4f'%(d['n_gpus'], d['value'], d['ms_per_step']))"
