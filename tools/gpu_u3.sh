#!/bin/bash
mkdir -p gpurun_out
N=16 DEPTH=40 ALLROWS=1 timeout 400 python tools/mem_cfg5.py > gpurun_out/u3_mem_n16.log 2>&1; tail -30 gpurun_out/u3_mem_n16.log | cut -c1-250
