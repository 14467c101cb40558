#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python bench_configs.py --configs 5 --qubit-scale 0.16 --depth-scale 0.7 --profile gpurun_out/u6_prof_cfg5.txt ) > gpurun_out/u6_cfg5.log 2>&1
cut -c1-60,100-200 gpurun_out/u6_prof_cfg5.txt | head -28; grep updates_per gpurun_out/u6_cfg5.log | cut -c1-300
( time timeout 300 python bench_configs.py --configs 3 --qubit-scale 0.32 --depth-scale 0.7 --profile gpurun_out/u6_prof_cfg3.txt ) > gpurun_out/u6_cfg3.log 2>&1
cut -c1-60,100-200 gpurun_out/u6_prof_cfg3.txt | head -28; grep updates_per gpurun_out/u6_cfg3.log | cut -c1-300
