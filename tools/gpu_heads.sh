#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bprmf.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r_pytest.log
timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
tail -1 gpurun_out/r_bench.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('ms %.4f e2e %.4f apply %.4f frac %.3f plan %.4f fused %.4f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['kernels']['plan_items']['ms'], d['kernels']['fused_score_loss_bwd']['ms']))"

