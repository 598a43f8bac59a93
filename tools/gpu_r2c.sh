#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2c_pytest.log)"
grep -E "^FAILED|^ERROR" gpurun_out/r2c_pytest.log | head -20
timeout 300 python tools/diag_dq.py 2>&1 | tail -6
B2R_FUSED=v6 timeout 300 python tools/diag_dq.py 2>&1 | tail -5
/usr/bin/time -v timeout 1200 python bench.py --steps 200 --warmup 10 > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
grep -E "Elapsed|Maximum resident" gpurun_out/r2c_bench.err
tail -1 gpurun_out/r2c_bench.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('c2 ms %.4f e2e %.4f roofline %s step_roofline %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['step_roofline']['frac']))
print('kernels', {k:v['ms'] for k,v in d['kernels'].items()})
print('contract', d['contract_route']['ms_per_step'], d['contract_route']['with_runner_shuffle']['ms_per_step'])
print('self_check', d['self_check'])
print('cpu', d.get('cpu_baseline'))
for k,v in d.get('workloads',{}).items():
    print(k, {a:v.get(a) for a in ('value','ms_per_step','epoch_s','error')}, 'e2e', v.get('e2e',{}).get('ms_per_step'), 'roof', v.get('roofline',{}).get('frac'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), (v.get('cpu_baseline') or {}).get('ms_per_step'))
print('clocks', d['clocks'])
"
tail -5 gpurun_out/r2c_bench.err
