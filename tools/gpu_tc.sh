#!/bin/bash
mkdir -p gpurun_out
timeout 240 python tools/tc_check.py > gpurun_out/m_tc.log 2>&1; echo "tc rc=$?" >> gpurun_out/m_tc.log
tail -10 gpurun_out/m_tc.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 500 --csv --log-file gpurun_out/m_sasrec_launches.csv \
    python tools/model_bench.py --steps 3 --warmup 2 > gpurun_out/m_ncu_models.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/m_sasrec_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault(r[ki][:70],[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:22]:
    print(f"{sum(v)/tot*100:5.1f}%  n={len(v):3d} avg={sum(v)/len(v)/1000:8.1f} us  {k}")
PY
