#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_gpu_bprmf.py::test_full_size_config2_properties > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_pytest.log
grep -E "^(FAILED|ERROR|E  )|passed|failed" gpurun_out/b_pytest.log | head -80
