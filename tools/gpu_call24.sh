#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/prof_outliers.py 30 2>&1 | grep -v Warn > gpurun_out/c24_b8.log
awk '{print $3}' gpurun_out/c24_b8.log | grep -E '^[0-9.]+$' | sort -n | awk '{a[NR]=$1} END{print "b8: min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
tail -1 gpurun_out/c24_b8.log | cut -c1-150
MPDO_JACOBI_B=16 MPDO_JACOBI_G=16 timeout 300 python tools/prof_outliers.py 30 2>&1 | grep -v Warn > gpurun_out/c24_b16.log
awk '{print $3}' gpurun_out/c24_b16.log | grep -E '^[0-9.]+$' | sort -n | awk '{a[NR]=$1} END{print "b16: min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
tail -1 gpurun_out/c24_b16.log | cut -c1-150
timeout 200 python bench_configs.py --configs 3 --qubit-scale 0.3 --depth-scale 0.4 2>&1 | grep '^{"config' | cut -c1-200
MPDO_JACOBI_B=16 timeout 200 python bench_configs.py --configs 3 --qubit-scale 0.3 --depth-scale 0.4 2>&1 | grep '^{"config' | cut -c1-200
