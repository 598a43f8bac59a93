#!/bin/bash
# round-end validation: full GPU test suite, smoke(), the default bench line, the reference arm, launch lists
mkdir -p gpurun_out
S=$SECONDS
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/final_pytest.log) [$((SECONDS-S)) s]"
grep -E "^FAILED|^ERROR" gpurun_out/final_pytest.log | head -20
S=$SECONDS
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
S=$SECONDS
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$? [$((SECONDS-S)) s]"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1])
print('c2 value %.4g %s  ms/step %.4f  e2e %.4g  roofline frac %s  step_roofline %s' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d.get('step_roofline', {}).get('frac')))
for k, v in (d.get('workloads') or {}).items():
    if isinstance(v, dict):
        print(' ', k, {kk: v[kk] for kk in ('ms_per_step', 'eager_ms_per_step', 'graph_error', 'epoch_s', 'value') if kk in v})
print('clocks', d.get('clocks'))
PY
S=$SECONDS
timeout 600 python bench.py --impl reference --steps 4 --warmup 0 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench_reference.err; echo "reference arm rc=$? [$((SECONDS-S)) s] $(cut -c1-300 gpurun_out/final_bench_reference.json | tail -1)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/final_launches_c4.csv python bench.py --workload c4 --steps 3 --warmup 3 --no_cpu_baseline > /dev/null 2>&1; echo "ncu c4 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c2.csv python bench.py --steps 20 --warmup 3 --headline_only --no_cpu_baseline > /dev/null 2>&1; echo "ncu c2 rc=$?"
