#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'pipe' --launch-skip 2 --launch-count 2 -o gpurun_out/r2t_pipe -f python tools/prof_dense.py > gpurun_out/r2t_ncu.log 2>&1
tail -3 gpurun_out/r2t_ncu.log
