#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tools/prof_outliers.py 200 2>&1 | grep -v Warn > gpurun_out/c12_outliers.log
grep -A15 OUTLIER gpurun_out/c12_outliers.log | head -70
awk '{print $3}' gpurun_out/c12_outliers.log | grep -E '^[0-9.]+$' | sort -n | awk '{a[NR]=$1} END{print "min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
grep -c "num_device_alloc" gpurun_out/c12_outliers.log
grep "num_device_alloc\|pool_MB [1-9]" gpurun_out/c12_outliers.log | head -20
