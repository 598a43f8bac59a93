#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --deselect tests/test_gpu_bprmf.py::test_full_size_config2_properties > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/c_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/c_pytest.log | head -20
for rpi in 0 1 2 4; do
  B2R_SEG_RPI=$rpi timeout 300 python bench.py --steps 300 --warmup 10 --no_cpu_baseline > gpurun_out/c_bench_rpi$rpi.json 2> gpurun_out/c_bench_rpi$rpi.err
  echo "RPI=$rpi"; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/c_bench_rpi$rpi.json'))
    print(' value %.3e ms/step %.4f e2e %.3e launches %d'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
    print(' kernels', {k:v['ms'] for k,v in d['kernels'].items()})
    print(' roofline', d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'], 'clocks', d['clocks'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/c_bench_rpi$rpi.err').read()[-1500:])
PY
done
