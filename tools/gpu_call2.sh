#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c2_pytest.log 2>&1
tail -3 gpurun_out/c2_pytest.log
echo "== DMMA"; timeout 120 python tools/bench_contract.py 2>&1 | tee gpurun_out/c2_contract_dmma.log
echo "== scalar"; MPDO_NO_DMMA=1 timeout 120 python tools/bench_contract.py 2>&1 | tee gpurun_out/c2_contract_scalar.log
( time timeout 400 python bench.py --no-cpu-baseline ) > gpurun_out/c2_bench.log 2>&1
tail -1 gpurun_out/c2_bench.log | cut -c1-300
( time timeout 200 python bench_configs.py --configs 5 --qubit-scale 0.16 --depth-scale 0.3 --profile gpurun_out/c2_prof_cfg5.txt ) > gpurun_out/c2_cfg5.log 2>&1
tail -4 gpurun_out/c2_cfg5.log | cut -c1-300
