#!/bin/bash
mkdir -p gpurun_out
( time timeout 400 python bench.py ) > gpurun_out/g_bench.log 2>&1
grep '^{' gpurun_out/g_bench.log | cut -c1-200
( time timeout 200 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/g_bench_ref.log 2>&1
rm -f gpurun_out/g_configs.jsonl
( time timeout 600 python bench_configs.py --configs 1,2,3 --out gpurun_out/g_configs.jsonl ) > gpurun_out/g_configs.log 2>&1
cut -c1-220 gpurun_out/g_configs.jsonl
