#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_time.py 2>&1 | head -8 | tee gpurun_out/r2p_gemm_time.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_linear_fwd_tc|k_gemm' --launch-count 2 -o gpurun_out/r2p_lin -f python tools/prof_dense.py > gpurun_out/r2p_ncu.log 2>&1
tail -3 gpurun_out/r2p_ncu.log
