#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/n_pytest.log | head -20
timeout 600 python tools/shard_bench.py --n_items 100000000 --n_users 1000000 --emb 128 --B 4096 --K 255 --steps 20 --warmup 5 > gpurun_out/n_c5_1.log 2>&1; grep -E '^\{|Error' gpurun_out/n_c5_1.log | head -3
timeout 600 python tools/model_bench.py --steps 20 --warmup 3 > gpurun_out/n_models.log 2>&1; tail -2 gpurun_out/n_models.log
