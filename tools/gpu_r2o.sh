#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py tests/test_gpu_tc.py -q -x > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2o_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2o_pytest.log | head -20
timeout 300 python tools/attn_time.py 2>&1 | tee gpurun_out/r2o_attn_time.txt
timeout 300 python tools/gemm_time.py 2>&1 | tee gpurun_out/r2o_gemm_time.txt
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no_cpu_baseline 2>gpurun_out/r2o_c4.err | tail -1 > gpurun_out/r2o_c4.json
python -c "
import json; d=json.load(open('gpurun_out/r2o_c4.json')); print('c4 graphed ms %.4f eager %.4f e2e %.4f loss %s %s'%(d['ms_per_step'], d['eager_ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d.get('graph_error')))"
tail -3 gpurun_out/r2o_c4.err | cut -c1-300
