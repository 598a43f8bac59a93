#!/bin/bash
mkdir -p gpurun_out
for M in "" inline legacy; do
  B2R_STEP=$M timeout 900 python -m pytest tests/test_gpu_c2_step.py tests/test_gpu_bprmf.py tests/test_gpu_runner_fit.py tests/test_gpu_zz_fit_golden.py tests/test_gpu_shard.py -q -x > gpurun_out/r2e_pytest_$M.log 2>&1; echo "pytest B2R_STEP='$M' rc=$? $(tail -1 gpurun_out/r2e_pytest_$M.log)"
  grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2e_pytest_$M.log | head -8
done
timeout 600 python -m pytest tests/test_gpu_optin_modes.py "tests/test_gpu_overlay.py::test_out_of_range_id_raises_like_the_reference" "tests/test_gpu_overlay.py::test_deep_models_through_unchanged_main_track_reference_cpu_run" -q > gpurun_out/r2e_pytest_misc.log 2>&1; echo "pytest misc rc=$? $(tail -1 gpurun_out/r2e_pytest_misc.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2e_pytest_misc.log | head -12
DIAG_FLUSH=1 timeout 120 python tools/diag_exact_adam.py 0.0 2>&1 | tail -30
pr() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$1: ms %.4f e2e %.4f | '%(d['ms_per_step'], d['e2e']['ms_per_step']) + ' '.join('%s %.4f'%(a,b['ms']) for a,b in k.items()) + ' | frac %s step_frac %s launches %s'%(d['roofline']['frac'], d['step_roofline']['frac'], d['gpu_launches']))"; }
for M in "" inline legacy; do
  B2R_STEP=$M timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | pr "step mode '$M' flash 223"
done
B2R_FLASH=143 timeout 300 python bench.py --steps 1000 --warmup 20 --no_cpu_baseline --headline_only 2>/dev/null | tail -1 | pr "step mode '' flash 143"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 60 --csv --log-file gpurun_out/r2e_launches.csv \
    python bench.py --steps 12 --warmup 8 --no_cpu_baseline --headline_only > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2e_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(',','')))
    except: pass
for k,v in agg.items(): print(f"  n={len(v):3d} avg={sum(v)/len(v)/1000:8.2f} us  {k}")
PY
