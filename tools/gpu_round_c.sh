#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_bprmf.py::test_full_size_config2_properties > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/f_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/f_pytest.log | head -20
grep -E "^E  " gpurun_out/f_pytest.log | head -30
for spi in 2 1; do
  B2R_FUSED_SPI=$spi timeout 300 python bench.py --steps 400 --warmup 10 --no_cpu_baseline > gpurun_out/f_bench_spi$spi.json 2> gpurun_out/f_bench_spi$spi.err
  echo "SPI=$spi"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/f_bench_spi$spi.json'))
    print(' value %.3e ms/step %.4f e2e %.3e (%.4f ms) launches %d'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches']))
    print(' kernels', {k:v['ms'] for k,v in d['kernels'].items()})
    print(' roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], 'clocks', d['clocks'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/f_bench_spi$spi.err').read()[-2500:])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 100 -c 150 --csv --log-file gpurun_out/f_launches.csv \
    python bench.py --steps 12 --warmup 8 --no_cpu_baseline > gpurun_out/f_ncu_list.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/f_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name'); gi=hdr.index('Grid Size')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault((r[ki][:60]+' grid'+r[gi], r[mi]),[]).append(v)
for (k,m),v in agg.items():
    if m=='gpu__time_duration.sum': print(f"  n={len(v):3d} avg={sum(v)/len(v)/1000:8.2f} us  {k}")
for (k,m),v in agg.items():
    if m!='gpu__time_duration.sum' and ('fused' in k or 'bucket_apply<' in k): print(f"  {m} avg={sum(v)/len(v)/1e6:.1f} MB  {k[:60]}")
PY
