#!/bin/bash
mkdir -p gpurun_out
for rpi in 1 4; do
B2R_SEG_RPI=$rpi timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 200 --csv --log-file gpurun_out/d_launches_rpi$rpi.csv \
    python bench.py --steps 12 --warmup 8 --no_cpu_baseline > gpurun_out/d_ncu_list_rpi$rpi.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/d_launches_rpi$rpi.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault((r[ki][:70], r[mi]),[]).append(v)
print("RPI=$rpi")
for (k,m),v in agg.items():
    if m=='gpu__time_duration.sum': print(f"  n={len(v):3d} avg={sum(v)/len(v)/1000:8.2f} us  {k}")
for (k,m),v in agg.items():
    if m!='gpu__time_duration.sum' and ('fused' in k or 'segment' in k): print(f"  {m} avg={sum(v)/len(v):.1f}  {k[:40]}")
PY
done
