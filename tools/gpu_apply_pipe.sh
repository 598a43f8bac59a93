#!/bin/bash
# k_apply_sorted pipeline experiment: correctness with the cp.async ring, bench per depth, one full ncu capture each.
mkdir -p gpurun_out
B2R_APPLY_PIPE=3 timeout 600 python -m pytest tests/test_gpu_bprmf.py tests/test_gpu_fullsize.py tests/test_gpu_runner_fit.py tests/test_gpu_eval.py -x -q > gpurun_out/p_pytest_pipe3.log 2>&1; echo "pytest(pipe3) rc=$?"; tail -3 gpurun_out/p_pytest_pipe3.log
for cfg in "0 0" "2 0" "3 0" "4 0" "6 0" "2 8" "3 2"; do
  set -- $cfg
  B2R_APPLY_PIPE=$1 B2R_APPLY_CTAS=$2 timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline > gpurun_out/p_bench_$1_$2.json 2> gpurun_out/p_bench_$1_$2.err
  tail -1 gpurun_out/p_bench_$1_$2.json | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('pipe $1 ctas $2: ms %.4f apply %.4f frac %.3f plan %.4f fused %.4f'%(d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['kernels']['plan_items']['ms'], d['kernels']['fused_score_loss_bwd']['ms']))"
done
for P in 0 3; do
B2R_APPLY_PIPE=$P timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_apply_sorted -s 6 -c 2 -o gpurun_out/p_apply_pipe$P -f python bench.py --steps 6 --warmup 3 --no_cpu_baseline > gpurun_out/p_ncu_$P.log 2>&1; echo "ncu($P) rc=$?"
done
