#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/u4_configs.jsonl
( time timeout 900 python bench_configs.py --configs 5,3 --out gpurun_out/u4_configs.jsonl ) > gpurun_out/u4_configs.log 2>&1
cut -c1-420 gpurun_out/u4_configs.jsonl; tail -4 gpurun_out/u4_configs.log | cut -c1-300
