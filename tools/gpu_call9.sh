#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_outliers.py 40 2>&1 | grep -v Warn | tee gpurun_out/c9_outliers.log | awk '{ if ($3+0 > 260) print }'
