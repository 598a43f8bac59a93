#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_models.py tests/test_gpu_fullsize.py tests/test_gpu_tc.py tests/test_gpu_optin_modes.py tests/test_gpu_zz_fit_golden.py -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2j_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2j_pytest.log | head -20
for LV in 1 0; do
B2R_SASREC_LIVE=$LV timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no_cpu_baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c4 LIVE=$LV ms %.4f e2e %.4f loss %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss']))"
done
timeout 600 python bench.py --steps 300 --warmup 10 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c2 ms %.4f contract %s exact %s'%(d['ms_per_step'], d['contract_route']['ms_per_step'], d['contract_route'].get('exact_adam')))"
