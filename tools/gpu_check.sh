#!/bin/bash
# Quick GPU check: parity tests, contraction rates, one bench run.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 100 python tools/bench_contract.py 2>&1 | tee gpurun_out/i_contract.log | head -5
timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/i_bench.json
python -c "
import json; d=json.load(open('gpurun_out/i_bench.json')); print(round(d['value'],1), d['ms_each_step'], 'e2e', round(d['e2e']['value'],1))"
