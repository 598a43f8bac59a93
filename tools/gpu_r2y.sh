#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_graph.py -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r2y_pytest.log)"
grep -E "^FAILED|^ERROR|^E  " gpurun_out/r2y_pytest.log | head -20
