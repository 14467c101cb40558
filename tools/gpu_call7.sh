#!/bin/bash
mkdir -p gpurun_out
nproc; uptime
for i in 1 2 3; do
  timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/c7_dmma_$i.json
  MPDO_NO_DMMA=1 timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/c7_scalar_$i.json
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c7_*.json')):
    d=json.loads(open(f).read())
    print(f, round(d['value'],1), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],1), 'contract_s', round(d['roofline']['kernel_seconds'],3))
PY
uptime
