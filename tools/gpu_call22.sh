#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_outliers.py 100 2>&1 | grep -v Warn > gpurun_out/c22_fused.log
awk '{print $3}' gpurun_out/c22_fused.log | grep -E '^[0-9.]+$' | sort -n | awk '{a[NR]=$1} END{print "fused: min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
grep -c num_device_alloc gpurun_out/c22_fused.log
for i in 1 2 3 4 5; do timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/c22_bench_$i.json; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c22_bench_*.json')):
    d=json.loads(open(f).read())
    print(f, round(d['value'],1), d['ms_each_step'], 'e2e', round(d['e2e']['value'],1))
PY
