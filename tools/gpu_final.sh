#!/bin/bash
# round-end evidence: full GPU test suite, smoke, the default bench line (with cpu_baseline), the reference arm,
# the ncu launch list of the bench command, one --set full capture of the two hot kernels, eval + model benches
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/z_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/z_pytest.log; tail -2 gpurun_out/z_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/z_smoke.log
timeout 400 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/z_bench.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; echo "ref rc=$?"; tail -c 700 gpurun_out/z_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 300 --csv --log-file gpurun_out/z_launches.csv \
    python bench.py --steps 12 --warmup 8 --no_cpu_baseline > gpurun_out/z_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_apply_sorted|k_bprmf_fused' -s 9 -c 3 \
    -o gpurun_out/z_prof -f python bench.py --steps 6 --warmup 3 --no_cpu_baseline > gpurun_out/z_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 200 python tools/eval_bench.py > gpurun_out/z_eval.log 2>&1; echo "eval rc=$?"; tail -2 gpurun_out/z_eval.log
timeout 400 python tools/model_bench.py > gpurun_out/z_models.log 2>&1; echo "models rc=$?"; grep -E '^\{' gpurun_out/z_models.log | cut -c1-300
