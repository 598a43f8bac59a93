#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_graph.py -q -x > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2m_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2m_pytest.log | head -20
for W in c3 c4; do
timeout 600 python bench.py --workload $W --steps 40 --warmup 3 --no_cpu_baseline 2>gpurun_out/r2m_$W.err | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$W graphed ms %.4f eager %.4f e2e %.4f loss %s'%(d['ms_per_step'], d['eager_ms_per_step'], d['e2e']['ms_per_step'], d['final_loss']), d.get('graph_error'))"
tail -3 gpurun_out/r2m_$W.err | cut -c1-300
done
