#!/bin/bash
mkdir -p gpurun_out
echo "== DMMA 16 warps"; timeout 120 python tools/bench_contract.py 2>&1 | tee gpurun_out/c17_contract_16w.log
echo "== DMMA 8 warps"; MPDO_DMMA_8WARPS=1 timeout 120 python tools/bench_contract.py 2>&1 | tee gpurun_out/c17_contract_8w.log
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/c17_bench_$i.json
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c17_bench_*.json')):
    d=json.loads(open(f).read())
    print(f, round(d['value'],1), d['ms_each_step'], 'e2e', round(d['e2e']['value'],1), 'contract_s', round(d['roofline']['kernel_seconds'],3), d['roofline']['largest_launch'])
PY
