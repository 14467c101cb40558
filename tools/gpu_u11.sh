#!/bin/bash
mkdir -p gpurun_out
SKIP_BIG_CLASSIC=1 timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u11_p2p.log 2>&1; tail -11 gpurun_out/u11_p2p.log
MPDO_JACOBI_NOP2P=1 SKIP_BIG_CLASSIC=1 timeout 300 python tools/bench_jacobi_big.py > gpurun_out/u11_nop2p.log 2>&1; tail -11 gpurun_out/u11_nop2p.log
