#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_c2_step.py tests/test_gpu_bprmf.py tests/test_gpu_optin_modes.py tests/test_gpu_zz_fit_golden.py tests/test_gpu_runner_fit.py "tests/test_gpu_overlay.py::test_out_of_range_id_raises_like_the_reference" "tests/test_gpu_overlay.py::test_bprmf_fused_whole_step_route_through_unchanged_main_equals_reference_sgd" -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest (direct step) rc=$? $(tail -1 gpurun_out/r2d_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2d_pytest.log | head -20
B2R_STEP=prefetch timeout 600 python -m pytest tests/test_gpu_c2_step.py tests/test_gpu_bprmf.py -q -x > gpurun_out/r2d_pytest_prefetch.log 2>&1; echo "pytest (prefetch step) rc=$? $(tail -1 gpurun_out/r2d_pytest_prefetch.log)"
timeout 120 python tools/diag_exact_adam.py 0.0 2>&1 | tail -24
pr() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$1: ms %.4f e2e %.4f | '%(d['ms_per_step'], d['e2e']['ms_per_step']) + ' '.join('%s %.4f'%(a,b['ms']) for a,b in k.items()) + ' | frac %s step_frac %s launches %s'%(d['roofline']['frac'], d['step_roofline']['frac'], d['gpu_launches']))"; }
for V in 223 143 243; do
  B2R_FLASH=$V timeout 300 python bench.py --steps 600 --warmup 20 --no_cpu_baseline --headline_only 2>gpurun_out/r2d_bench_$V.err | tail -1 > gpurun_out/r2d_bench_$V.json; pr "direct flash $V" < gpurun_out/r2d_bench_$V.json
done
B2R_STEP=prefetch timeout 300 python bench.py --steps 600 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | pr "prefetch flash 223"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/r2d_launches.csv \
    python bench.py --steps 12 --warmup 8 --no_cpu_baseline --headline_only > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2d_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(',','')))
    except: pass
for k,v in agg.items(): print(f"  n={len(v):3d} avg={sum(v)/len(v)/1000:8.2f} us  {k}")
PY
SECONDS=0
timeout 1500 python bench.py --steps 200 --warmup 10 > gpurun_out/r2d_bench_full.json 2> gpurun_out/r2d_bench_full.err; echo "full bench rc=$? in ${SECONDS}s"
tail -1 gpurun_out/r2d_bench_full.json | python -c "
import sys,json
d=json.loads(sys.stdin.read())
print('c2 ms %.4f e2e %.4f roofline %s step_roofline %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['step_roofline']['frac']))
print('contract', d['contract_route']['ms_per_step'], d['contract_route']['with_runner_shuffle']['ms_per_step'])
print('self_check', d['self_check'])
print('cpu', d.get('cpu_baseline'))
for k,v in d.get('workloads',{}).items():
    print(k, {a:v.get(a) for a in ('value','ms_per_step','epoch_s','error')}, 'e2e', v.get('e2e',{}).get('ms_per_step'), 'roof', v.get('roofline',{}).get('frac'), 'cpu', (v.get('cpu_baseline') or {}).get('value'), (v.get('cpu_baseline') or {}).get('ms_per_step'), (v.get('cpu_baseline') or {}).get('error'))
print('clocks', d['clocks'])
"
tail -3 gpurun_out/r2d_bench_full.err
