#!/bin/bash
mkdir -p gpurun_out
SKIP_BIG_CLASSIC=1 timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u12_tree.log 2>&1; tail -11 gpurun_out/u12_tree.log
timeout 400 python tools/prof_cfg4.py > gpurun_out/u12_cfg4.log 2>&1
grep -v "^---" gpurun_out/u12_cfg4.log | awk '{print substr($0,1,76) "  |  " substr($0,length($0)-44)}' | head -40
