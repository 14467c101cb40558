#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/mem_cfg5.py > gpurun_out/u2_mem_default.log 2>&1; tail -14 gpurun_out/u2_mem_default.log | cut -c1-250
MPDO_GROUPING=0 timeout 300 python tools/mem_cfg5.py > gpurun_out/u2_mem_nogroup.log 2>&1; tail -14 gpurun_out/u2_mem_nogroup.log | cut -c1-250
MPDO_ENV_SWEEP=0 timeout 300 python tools/mem_cfg5.py > gpurun_out/u2_mem_noenv.log 2>&1; tail -14 gpurun_out/u2_mem_noenv.log | cut -c1-250
MPDO_GROUPING=0 MPDO_ENV_SWEEP=0 timeout 300 python tools/mem_cfg5.py > gpurun_out/u2_mem_neither.log 2>&1; tail -14 gpurun_out/u2_mem_neither.log | cut -c1-250
