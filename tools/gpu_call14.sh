#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/c14_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/c14_pytest.log | tail -3
for v in track notrack; do
  if [ $v = notrack ]; then export MPDO_JACOBI_NOTRACK=1; fi
  timeout 300 python tools/prof_outliers.py 40 2>&1 | grep -v Warn > gpurun_out/c14_$v.log
  awk '{print $3}' gpurun_out/c14_$v.log | grep -E '^[0-9.]+$' | sort -n | awk -v v=$v '{a[NR]=$1} END{print v, "min",a[1],"med",a[int(NR/2)],"p90",a[int(NR*0.9)],"max",a[NR], NR}'
  tail -1 gpurun_out/c14_$v.log | cut -c1-200
  timeout 200 python bench_configs.py --configs 3 --qubit-scale 0.3 --depth-scale 0.4 2>&1 | grep '^{"config' | cut -c1-260
done
