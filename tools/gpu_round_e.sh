#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/j_pytest.log | head -20
grep -E "^E  " gpurun_out/j_pytest.log | head -20
timeout 300 python bench.py --steps 400 --warmup 10 --no_cpu_baseline > gpurun_out/j_bench.json 2> gpurun_out/j_bench.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/j_bench.json'))
    print(' value %.3e ms/step %.4f e2e %.3e (%.4f ms) launches %d'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
    print(' kernels', {k:v['ms'] for k,v in d['kernels'].items()})
    print(' roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/j_bench.err').read()[-2500:])
PY
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 130 --csv --log-file gpurun_out/j_launches.csv \
    python bench.py --steps 10 --warmup 8 --no_cpu_baseline > gpurun_out/j_ncu_list.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/j_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault((r[ki][:44]+' grid'+r[gi], r[mi]),[]).append(v)
for (k,m),v in agg.items():
    if m=='gpu__time_duration.sum':
        n=len(v)
        print(f"  n={n:3d} avg={sum(v)/n/1000:8.2f} us  inst={sum(agg[(k,'smsp__inst_executed.sum')])/n/1e6:7.2f}M  dram r/w={sum(agg[(k,'dram__bytes_read.sum')])/n/1e6:6.1f}/{sum(agg[(k,'dram__bytes_write.sum')])/n/1e6:6.1f} MB  {k}")
PY
timeout 600 python tools/model_bench.py --steps 20 --warmup 3 > gpurun_out/j_models.log 2>&1; tail -3 gpurun_out/j_models.log
