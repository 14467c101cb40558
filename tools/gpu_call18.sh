#!/bin/bash
mkdir -p gpurun_out
echo "== cluster"; timeout 120 python tools/bench_chol.py 2>&1 | tee gpurun_out/c18_chol_cluster.log
echo "== global barrier"; MPDO_CHOL_NOCLUSTER=1 timeout 120 python tools/bench_chol.py 2>&1 | tee gpurun_out/c18_chol_global.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_outliers.py 60 2>&1 | grep -v Warn > gpurun_out/c18_outliers.log
awk '{print $3}' gpurun_out/c18_outliers.log | grep -E '^[0-9.]+$' | sort -n | awk '{a[NR]=$1} END{print "min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
tail -2 gpurun_out/c18_outliers.log | cut -c1-160
