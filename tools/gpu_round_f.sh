#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_fullsize.py > gpurun_out/l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/l_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/l_pytest.log | head -20
timeout 300 python bench.py --steps 400 --warmup 10 --no_cpu_baseline > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err
python -c "
import json; d=json.load(open('gpurun_out/l_bench.json')); print(' value %.3e ms/step %.4f e2e %.3e launches %d'%(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])); print({k:v['ms'] for k,v in d['kernels'].items()}); print(d['roofline'])" || tail -20 gpurun_out/l_bench.err
