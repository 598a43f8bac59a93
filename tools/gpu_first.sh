#!/bin/bash
# first GPU contact: parity tests, smoke, design probe, bench, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 120 python __graft_entry__.py > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 120 ./build/gather_bench > gpurun_out/a_gather_bench.txt 2>&1
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/a_launches.csv \
    python bench.py --steps 3 --warmup 3 --no_cpu_baseline > gpurun_out/a_ncu_list.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_rowdot_fwd|k_segment_apply|k_rowdot_bwd_query' -s 9 -c 3 \
    -o gpurun_out/a_prof python bench.py --steps 2 --warmup 3 --no_cpu_baseline > gpurun_out/a_ncu_full.log 2>&1
tail -5 gpurun_out/a_pytest.log; cat gpurun_out/a_smoke.log | tail -3; cat gpurun_out/a_gather_bench.txt; cat gpurun_out/a_bench.json; tail -3 gpurun_out/a_bench.err
