#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_listwise.py tests/test_gpu_optin_modes.py tests/test_gpu_shard.py -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2f_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2f_pytest.log | head -20
pr() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print(sys.argv[1] + ': ms %.4f e2e %.4f | '%(d['ms_per_step'], d['e2e']['ms_per_step']) + ' '.join('%s %.4f'%(a,b['ms']) for a,b in k.items()) + ' | frac %s step_frac %s launches %s'%(d['roofline']['frac'], d['step_roofline']['frac'], d['gpu_launches']))" "$1"; }
for PA in 0 1; do for FL in 223 143; do
  B2R_PLAN_AFTER=$PA B2R_FLASH=$FL timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | pr "plan_after=$PA flash=$FL"
done; done
B2R_STEP=inline timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | pr "inline"
B2R_STEP=legacy timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | pr "legacy"
timeout 600 python tools/shard_bench.py --n_items 100000000 --n_users 1000000 --emb 128 --B 4096 --K 255 --steps 20 --warmup 5 --torch_profile > gpurun_out/r2f_c5_n1_profile.log 2>&1
grep -E "^\{" gpurun_out/r2f_c5_n1_profile.log | cut -c1-160
grep -E "Name|k_|void|Memcpy|Memset|aten::" gpurun_out/r2f_c5_n1_profile.log | awk '{print}' | cut -c1-200 | head -40
