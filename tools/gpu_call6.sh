#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python bench.py --no-cpu-baseline ) > gpurun_out/c6_bench.log 2>&1
grep '^{"metric' gpurun_out/c6_bench.log | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"jacobi_persistent|chol_kernel" -s 2 -c 2 -f -o gpurun_out/c6_jacobi python tools/ncu_targets.py > gpurun_out/c6_ncu.log 2>&1
ls -la gpurun_out/c6*.ncu-rep
