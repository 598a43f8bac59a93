#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_shard.py -q -x > gpurun_out/o_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/o_pytest.log | head
timeout 900 python tools/shard_bench.py --n_items 100000000 --n_users 1000000 --emb 128 --B 4096 --K 255 --steps 10 --warmup 3 --torch_profile > gpurun_out/o_shard_prof.log 2>&1
grep -v "^-" gpurun_out/o_shard_prof.log | cut -c1-70,150-230 | head -16; grep '^{' gpurun_out/o_shard_prof.log | cut -c1-200
