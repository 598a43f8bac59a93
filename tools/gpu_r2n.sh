#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload c3 --steps 40 --warmup 3 --no_cpu_baseline >gpurun_out/r2n_c3.out 2>gpurun_out/r2n_c3.err
grep -n "Error\|File \"/root/repo\|File \"/tmp/code\|^    " gpurun_out/r2n_c3.err | cut -c1-260 | head -80
