#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 330 --csv --log-file gpurun_out/t_shard_launches.csv \
    python tools/shard_bench.py --n_items 100000000 --n_users 1000000 --emb 128 --B 4096 --K 255 --steps 4 --warmup 3 > gpurun_out/t_shard_ncu.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/t_shard_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: v=float(r[vi].replace(',',''))
    except: continue
    agg.setdefault(r[ki][:86],[]).append(v)
tot=sum(sum(v) for v in agg.values())
print("total us over captured launches: %.0f"%(tot/1000))
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1]))[:28]:
    print(f"{sum(v)/tot*100:5.1f}%  n={len(v):3d} avg={sum(v)/len(v)/1000:8.1f} us  {k}")
PY
tail -2 gpurun_out/t_shard_ncu.log
