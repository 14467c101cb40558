#!/bin/bash
mkdir -p gpurun_out
echo "== timing on, 12 workers"
OUTLIER_TIMING=1 timeout 300 python tools/prof_outliers.py 60 2>&1 | grep -v Warn | tee gpurun_out/c10_timing.log | awk '{ if ($3+0 > 330) print }'
echo "== 5 workers"
MPDO_STRAND_WORKERS=5 timeout 300 python tools/prof_outliers.py 60 2>&1 | grep -v Warn | tee gpurun_out/c10_w5.log | awk '{ if ($3+0 > 260) print }'
tail -2 gpurun_out/c10_timing.log; tail -2 gpurun_out/c10_w5.log
