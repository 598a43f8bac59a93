#!/bin/bash
# round 2, call b: parity of the flash kernel + full-size step test; A/B of its variants; ncu of the new kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2b_pytest.log)"
grep -E "FAILED|Error" gpurun_out/r2b_pytest.log | head -20
pr() { python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$1: ms %.4f e2e %.4f | '%(d['ms_per_step'], d['e2e']['ms_per_step']) + ' '.join('%s %.4f'%(a,b['ms']) for a,b in k.items()) + ' | loss %.6f'%d['final_loss'])"; }
for V in 43 44 26 28 42; do
  B2R_FLASH=$V timeout 300 python bench.py --steps 600 --warmup 20 --no_cpu_baseline 2>gpurun_out/r2b_bench_$V.err | tail -1 > gpurun_out/r2b_bench_$V.json; pr "flash $V" < gpurun_out/r2b_bench_$V.json
done
B2R_FUSED=v6 timeout 300 python bench.py --steps 600 --warmup 20 --no_cpu_baseline 2>/dev/null | tail -1 | pr "v6"
B2R_FUSED=v6 B2R_BUCKET_SORT=bitonic timeout 300 python bench.py --steps 600 --warmup 20 --no_cpu_baseline 2>/dev/null | tail -1 | pr "v6+bitonic (round-1 kernels)"
# launch list (times alone) + full capture of the flash kernel and the counting sort
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 160 --csv --log-file gpurun_out/r2b_launches.csv \
    python bench.py --steps 12 --warmup 8 --no_cpu_baseline > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2b_launches.csv')) if len(r)>10]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    try: agg.setdefault(r[ki][:70],[]).append(float(r[vi].replace(',','')))
    except: pass
for k,v in agg.items(): print(f"  n={len(v):3d} avg={sum(v)/len(v)/1000:8.2f} us  {k}")
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bprmf_flash|k_bucket_sort' -s 8 -c 3 -o gpurun_out/r2b_flash \
    python bench.py --steps 12 --warmup 8 --no_cpu_baseline > gpurun_out/r2b_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/r2b_flash.ncu-rep
