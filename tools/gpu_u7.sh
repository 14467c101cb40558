#!/bin/bash
mkdir -p gpurun_out
MPDO_JACOBI_NOWIDE=1 timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u7_old.log 2>&1; cat gpurun_out/u7_old.log | tail -14
timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u7_wide.log 2>&1; cat gpurun_out/u7_wide.log | tail -14
