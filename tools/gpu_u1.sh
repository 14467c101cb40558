#!/bin/bash
# Session-4 check: GPU tests at HEAD, default bench (both arms), cfg5 and cfg3 at full size with the memory release.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/u1_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/u1_pytest.log | tail -2
( time timeout 400 python bench.py ) > gpurun_out/u1_bench.log 2>&1
grep '^{' gpurun_out/u1_bench.log | cut -c1-200
( time timeout 300 python bench.py --impl reference ) > gpurun_out/u1_bench_ref.log 2>&1
grep '^{' gpurun_out/u1_bench_ref.log | cut -c1-200
rm -f gpurun_out/u1_configs.jsonl
( time timeout 700 python bench_configs.py --configs 5,3 --out gpurun_out/u1_configs.jsonl ) > gpurun_out/u1_configs.log 2>&1
cut -c1-330 gpurun_out/u1_configs.jsonl; tail -3 gpurun_out/u1_configs.log | cut -c1-300
