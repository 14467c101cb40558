#!/bin/bash
mkdir -p gpurun_out
export MPDO_JACOBI_NOWIDE=1
MPDO_TRACE=1 timeout 300 python bench_configs.py --configs 3 --qubit-scale 0.32 --depth-scale 0.7 --no-warmup 2>&1 | grep "jacobi n=" | sort | uniq -c | sort -k1,1nr > gpurun_out/u8_sizes_cfg3.txt
head -40 gpurun_out/u8_sizes_cfg3.txt
MPDO_TRACE=1 timeout 300 python bench_configs.py --configs 5 --qubit-scale 0.16 --depth-scale 0.7 --no-warmup 2>&1 | grep "jacobi n=" | sort | uniq -c | sort -k1,1nr > gpurun_out/u8_sizes_cfg5.txt
awk '{split($4,a,"="); n=a[2]; c[int(n/64)*64]+=$1} END {for (k in c) print k, c[k]}' gpurun_out/u8_sizes_cfg5.txt | sort -n
MPDO_JACOBI_B=4 SKIP_BIG_CLASSIC=1 timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u8_b4.log 2>&1; grep -v "n=1024 B=1 classic\|n= 768 B=1 classic" gpurun_out/u8_b4.log | tail -12
