#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/ko_time.py 2>&1 | tee gpurun_out/r2w_ko_time.txt
