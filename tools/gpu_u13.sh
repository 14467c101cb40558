#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_prims.py -x -q -k "eigh" 2>&1 | tail -3
LAYER=12 timeout 200 python tools/prof_critical.py > gpurun_out/u13_critical12.log 2>&1
cut -c1-230 gpurun_out/u13_critical12.log | head -90
