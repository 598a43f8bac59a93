#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_bprmf_fused|k_bucket_apply<' -s 20 -c 3 \
    -o gpurun_out/g_prof python bench.py --steps 6 --warmup 6 --no_cpu_baseline > gpurun_out/g_ncu_full.log 2>&1
ls -la gpurun_out/g_prof.ncu-rep
