#!/bin/bash
mkdir -p gpurun_out
SKIP_BIG_CLASSIC=1 timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u9_cols.log 2>&1; tail -14 gpurun_out/u9_cols.log
MPDO_JACOBI_COLS4=1 SKIP_BIG_CLASSIC=1 timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u9_cols4.log 2>&1; tail -14 gpurun_out/u9_cols4.log
