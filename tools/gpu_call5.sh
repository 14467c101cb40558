#!/bin/bash
mkdir -p gpurun_out
echo "== default"; timeout 120 python tools/bench_contract.py 2>&1 | tee gpurun_out/c5_contract_default.log
echo "== DMMA all"; MPDO_DMMA_ALL=1 timeout 120 python tools/bench_contract.py 2>&1 | head -4 | tee gpurun_out/c5_contract_dmma.log
echo "== scalar"; MPDO_NO_DMMA=1 timeout 120 python tools/bench_contract.py 2>&1 | head -4 | tee gpurun_out/c5_contract_scalar.log
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c5_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/c5_pytest.log | tail -3
( time timeout 400 python bench.py --no-cpu-baseline ) > gpurun_out/c5_bench.log 2>&1
tail -5 gpurun_out/c5_bench.log | cut -c1-200
