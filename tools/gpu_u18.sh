#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/w_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/w_pytest.log | tail -2
grep -h "cfg3\|cfg5" gpurun_out/parity_big.jsonl | tail -4 | cut -c1-330
( time timeout 300 python bench_configs.py --configs 3 --qubit-scale 0.32 --depth-scale 0.7 ) > gpurun_out/w_cfg3.log 2>&1
grep updates_per gpurun_out/w_cfg3.log | cut -c1-330
MPDO_EIGH_SHIFT_MIN=0 timeout 300 python bench_configs.py --configs 3 --qubit-scale 0.32 --depth-scale 0.7 2>&1 | grep updates_per | cut -c1-330
