#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 100 python tools/bench_chol.py 2>&1 | grep eigh
echo "== G=32"; MPDO_JACOBI_G=32 timeout 100 python tools/bench_chol.py 2>&1 | grep eigh
echo "== B=8"; MPDO_JACOBI_B=8 timeout 100 python tools/bench_chol.py 2>&1 | grep eigh
echo "== B=8 G=32"; MPDO_JACOBI_B=8 MPDO_JACOBI_G=32 timeout 100 python tools/bench_chol.py 2>&1 | grep eigh
