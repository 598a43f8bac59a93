#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/tc_check.py 2>&1 | tail -14 | cut -c1-220 | tee gpurun_out/r2q_tc_check.txt
timeout 300 python tools/gemm_time.py 2>&1 | head -7 | tee gpurun_out/r2q_gemm_time.txt
for TC in 0 1; do
B2R_TC_LINEAR=$TC timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no_cpu_baseline 2>gpurun_out/r2q_c4_$TC.err | tail -1 > gpurun_out/r2q_c4_$TC.json
python -c "
import json; d=json.load(open('gpurun_out/r2q_c4_$TC.json')); print('TC_LINEAR=$TC c4 graphed ms %.4f eager %.4f e2e %.4f loss %s %s'%(d['ms_per_step'], d['eager_ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d.get('graph_error')))"
done
