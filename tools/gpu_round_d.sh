#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_bprmf.py::test_full_size_config2_properties > gpurun_out/i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/i_pytest.log | head -20
grep -E "^E  " gpurun_out/i_pytest.log | head -20
for rb in 2; do
  B2R_BUCKET_RB=$rb timeout 300 python bench.py --steps 400 --warmup 10 --no_cpu_baseline > gpurun_out/i_bench_rb$rb.json 2> gpurun_out/i_bench_rb$rb.err
  echo "RB=$rb"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/i_bench_rb$rb.json'))
    print(' value %.3e ms/step %.4f e2e %.3e (%.4f ms) launches %d'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
    print(' kernels', {k:v['ms'] for k,v in d['kernels'].items()})
    print(' roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/i_bench_rb$rb.err').read()[-2500:])
PY
  B2R_BUCKET_RB=$rb timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/i_launches_rb$rb.csv \
    python bench.py --steps 10 --warmup 8 --no_cpu_baseline > gpurun_out/i_ncu_list.log 2>&1
  python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/i_launches_rb$rb.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault((r[ki][:44]+' grid'+r[gi], r[mi]),[]).append(v)
for (k,m),v in agg.items():
    if m=='gpu__time_duration.sum': print(f"  n={len(v):3d} avg={sum(v)/len(v)/1000:8.2f} us  inst={sum(agg[(k,'smsp__inst_executed.sum')])/len(v)/1e6:7.2f}M  {k}")
PY
done
