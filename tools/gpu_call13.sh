#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c13_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/c13_pytest.log | tail -3
rm -f gpurun_out/c13_configs.jsonl
( time timeout 1500 python bench_configs.py --configs 1,2,3,4,5 --out gpurun_out/c13_configs.jsonl ) > gpurun_out/c13_configs.log 2>&1
cut -c1-330 gpurun_out/c13_configs.jsonl
