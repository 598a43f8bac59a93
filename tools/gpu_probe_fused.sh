#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bprmf.py -q -x -k "fused or c_side or fixture or prefetch" > gpurun_out/k_pytest.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/k_pytest.log | head -20
for cfg in "3 3" "3 6" "3 12" "2 6" "1 6"; do
  set -- $cfg
  B2R_FUSED_STOP=$1 B2R_FUSED_GRID=$2 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:k_bprmf_fused -s 10 -c 6 --csv --log-file gpurun_out/k_probe.csv \
    python bench.py --steps 10 --warmup 8 --no_cpu_baseline > gpurun_out/k_probe.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/k_probe.csv')) if len(r)>10]
hdr=rows[0]; vi=hdr.index('Metric Value'); mi=hdr.index('Metric Name')
t=[float(r[vi].replace(',','')) for r in rows[1:] if r[mi]=='gpu__time_duration.sum']
i=[float(r[vi].replace(',','')) for r in rows[1:] if r[mi]=='smsp__inst_executed.sum']
print("STOP=$1 GRID=$2x148: avg %.2f us, inst %.2fM" % (sum(t)/len(t)/1000, sum(i)/len(i)/1e6))
PY
done
timeout 300 python bench.py --steps 400 --warmup 10 --no_cpu_baseline > gpurun_out/k_bench.json 2> gpurun_out/k_bench.err
python -c "
import json; d=json.load(open('gpurun_out/k_bench.json')); print(' value %.3e ms/step %.4f e2e %.3e'%(d['value'], d['ms_per_step'], d['e2e']['value'])); print({k:v['ms'] for k,v in d['kernels'].items()})"
