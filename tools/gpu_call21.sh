#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/prof_outliers.py 30 2>&1 | grep -v Warn > gpurun_out/c21_fused.log
awk '{print $3}' gpurun_out/c21_fused.log | grep -E '^[0-9.]+$' | sort -n | awk '{a[NR]=$1} END{print "fused: min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
tail -1 gpurun_out/c21_fused.log | cut -c1-170
MPDO_NO_FUSE=1 timeout 300 python tools/prof_outliers.py 30 2>&1 | grep -v Warn > gpurun_out/c21_nofuse.log
awk '{print $3}' gpurun_out/c21_nofuse.log | grep -E '^[0-9.]+$' | sort -n | awk '{a[NR]=$1} END{print "nofuse: min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
tail -1 gpurun_out/c21_nofuse.log | cut -c1-170
for i in 1 2 3 4; do timeout 300 python bench.py --no-cpu-baseline 2>&1 | grep '^{"metric' > gpurun_out/c21_bench_$i.json; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/c21_bench_*.json')):
    d=json.loads(open(f).read())
    print(f, round(d['value'],1), d['ms_each_step'], 'e2e', round(d['e2e']['value'],1), d['config']['bond_dims_after_timed_region'])
PY
