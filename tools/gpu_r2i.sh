#!/bin/bash
mkdir -p gpurun_out
timeout 120 ./build/tma_gather4_probe 2>&1 | tail -8
SECONDS=0
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest -m gpu rc=$? in ${SECONDS}s: $(tail -1 gpurun_out/r2i_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2i_pytest.log | head -20
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SECONDS=0
timeout 1500 python bench.py --steps 200 --warmup 10 > gpurun_out/r2i_bench_full.json 2> gpurun_out/r2i_bench_full.err; echo "full bench rc=$? in ${SECONDS}s"
tail -1 gpurun_out/r2i_bench_full.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('c2 ms %.4f value %.4e e2e %.4f roofline %s step_roofline %.3f launches %s'%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['step_roofline']['frac'], d['gpu_launches']))
print('kernels', {k:v['ms'] for k,v in d['kernels'].items()})
print('contract', d['contract_route']['ms_per_step'], d['contract_route']['with_runner_shuffle']['ms_per_step'])
print('self_check', d['self_check']['ok'], d['self_check']['max_weight_err_well_conditioned'])
print('cpu', d.get('cpu_baseline',{}).get('value'), d.get('cpu_baseline',{}).get('ms_per_step'), d.get('cpu_baseline',{}).get('kind'), d.get('cpu_baseline',{}).get('cores'))
for k,v in d.get('workloads',{}).items():
    print(k, {a:v.get(a) for a in ('value','ms_per_step','epoch_s','error')}, 'e2e', v.get('e2e',{}).get('ms_per_step'), 'roof', v.get('roofline',{}).get('frac'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), (v.get('cpu_baseline') or {}).get('ms_per_step'), (v.get('cpu_baseline') or {}).get('error'))
print('clocks', d['clocks'])
"
SECONDS=0
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 2>/dev/null | tail -1 | cut -c1-400; echo "reference arm in ${SECONDS}s"
for SP in high low; do
B2R_SIDE_PRIO=$SP timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('side prio $SP: ms %.4f e2e %.4f | '%(d['ms_per_step'], d['e2e']['ms_per_step']) + ' '.join('%s %.4f'%(a,b['ms']) for a,b in k.items()))"
done
for W in 7 8; do for FL in 143 223; do
B2R_FLASH_WARPS=$W B2R_FLASH=$FL timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('flash warps $W variant $FL: ms %.4f e2e %.4f | '%(d['ms_per_step'], d['e2e']['ms_per_step']) + ' '.join('%s %.4f'%(a,b['ms']) for a,b in k.items()))"
done; done
