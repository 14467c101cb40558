#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_prims.py -x -q -k "blocked_factorisation" 2>&1 | tail -15 | cut -c1-200
SKIP_BIG_CLASSIC=1 timeout 200 python tools/bench_jacobi_big.py 2>&1 | grep "chol+jacobi" > gpurun_out/u21_panel.log; cat gpurun_out/u21_panel.log
