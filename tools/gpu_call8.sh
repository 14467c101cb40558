#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4; do
  timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/c8_warm_$i.json
  MPDO_SCRATCH_PREWARM_MB=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/c8_cold_$i.json
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c8_*.json')):
    d=json.loads(open(f).read())
    print(f, round(d['value'],1), d['ms_each_step'], 'e2e', round(d['e2e']['value'],1), d['config']['scratch_pool_GiB'])
PY
