#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/shard_bench.py --n_items 100000000 --n_users 1000000 --emb 128 --B 4096 --K 255 --steps 10 --warmup 3 --torch_profile > gpurun_out/t_shard_prof.log 2>&1
grep -v "^-" gpurun_out/t_shard_prof.log | cut -c1-200 | head -45
