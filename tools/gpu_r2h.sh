#!/bin/bash
mkdir -p gpurun_out
DIAG_FLUSH=1 timeout 120 python tools/diag_exact_adam.py 0.0 2>&1 | grep -v "rows off" | tail -95
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_dense.py tests/test_gpu_models.py tests/test_gpu_fullsize.py -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2h_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2h_pytest.log | head -20
timeout 600 python bench.py --workload c4 --steps 20 --warmup 3 --no_cpu_baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('c4 ms %.4f e2e %.4f loss %s tensor %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['final_loss'], d['tensor_roofline']['achieved']))"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2h_c4_launches.csv python bench.py --workload c4 --steps 3 --warmup 3 --no_cpu_baseline > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2h_c4_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(',','')))
    except: pass
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:24]: print(f"  n={len(v):4d} avg={sum(v)/len(v)/1000:8.2f} us tot={sum(v)/1e3:9.1f} {100*sum(v)/tot:5.1f}%  {k}")
PY
