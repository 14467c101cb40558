#!/bin/bash
# Both bench arms as the driver runs them at N = 1, outputs kept under gpurun_out/.
mkdir -p gpurun_out
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/h_bench_ref.log 2>&1
( time timeout 400 python bench.py ) > gpurun_out/h_bench.log 2>&1
grep '^{' gpurun_out/h_bench.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
