#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_dw_check.py 2>&1 | tail -14 | cut -c1-220 | tee gpurun_out/r2r_tc_dw_check.txt
timeout 300 python tools/gemm_time.py 2>&1 | head -7 | tee gpurun_out/r2r_gemm_time.txt
timeout 600 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py tests/test_gpu_tc.py tests/test_gpu_zz_fit_golden.py -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2r_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2r_pytest.log | head -20
for W in c4 c3; do
timeout 600 python bench.py --workload $W --steps 20 --warmup 3 --no_cpu_baseline 2>gpurun_out/r2r_$W.err | tail -1 > gpurun_out/r2r_$W.json
python -c "
import json; d=json.load(open('gpurun_out/r2r_$W.json')); print('$W graphed ms %.4f eager %.4f e2e %.4f loss %s %s'%(d['ms_per_step'], d['eager_ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d.get('graph_error')))"
done
