#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/c3_pytest.log | tail -3
( time timeout 300 python bench_configs.py --configs 4 --circuits 64 --chunk 64 --profile gpurun_out/c3_prof_cfg4.txt ) > gpurun_out/c3_cfg4.log 2>&1
grep -o '"seconds": [0-9.]*\|"circuits_per_s": [0-9.]*\|batch_vs_single[^,]*' gpurun_out/c3_cfg4.log
( time timeout 300 python bench_configs.py --configs 4 --circuits 128 --chunk 128 ) > gpurun_out/c3_cfg4_128.log 2>&1
grep -o '"seconds": [0-9.]*\|"circuits_per_s": [0-9.]*\|batch_vs_single[^,]*' gpurun_out/c3_cfg4_128.log
( time timeout 400 python bench.py --no-cpu-baseline ) > gpurun_out/c3_bench.log 2>&1
tail -5 gpurun_out/c3_bench.log | cut -c1-200
