#!/bin/bash
# Round-1 re-entry validation: GPU parity tests, the bench (both arms), and per-kernel device-time tables of scaled cfg3/4/5.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -3 gpurun_out/c1_pytest.log
( time timeout 400 python bench.py ) > gpurun_out/c1_bench.log 2>&1
tail -2 gpurun_out/c1_bench.log | cut -c1-600
( time timeout 200 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/c1_bench_ref.log 2>&1
( time timeout 200 python bench_configs.py --configs 5 --qubit-scale 0.16 --depth-scale 0.3 --profile gpurun_out/c1_prof_cfg5.txt ) > gpurun_out/c1_cfg5.log 2>&1
tail -1 gpurun_out/c1_cfg5.log | cut -c1-400
( time timeout 200 python bench_configs.py --configs 4 --circuits 64 --chunk 64 --profile gpurun_out/c1_prof_cfg4.txt ) > gpurun_out/c1_cfg4.log 2>&1
tail -1 gpurun_out/c1_cfg4.log | cut -c1-400
( time timeout 200 python bench_configs.py --configs 3 --qubit-scale 0.3 --depth-scale 0.4 --profile gpurun_out/c1_prof_cfg3.txt ) > gpurun_out/c1_cfg3.log 2>&1
tail -1 gpurun_out/c1_cfg3.log | cut -c1-400
