#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/u10_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/u10_pytest.log | tail -2
( time timeout 400 python bench_configs.py --configs 5 --qubit-scale 0.16 --depth-scale 0.7 ) > gpurun_out/u10_cfg5.log 2>&1
grep updates_per gpurun_out/u10_cfg5.log | cut -c1-330
( time timeout 300 python bench_configs.py --configs 3 --qubit-scale 0.32 --depth-scale 0.7 ) > gpurun_out/u10_cfg3.log 2>&1
grep updates_per gpurun_out/u10_cfg3.log | cut -c1-330
