#!/bin/bash
# Round-2 closing evidence (session 4, final code): GPU tests, both bench arms, ncu launch list of the timed region,
# full captures of the factorisation kernels incl. the new long-row Jacobi, and the five configurations at full size.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/z_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/z_pytest.log | tail -2
( time timeout 400 python bench.py ) > gpurun_out/z_bench.log 2>&1
grep '^{' gpurun_out/z_bench.log | cut -c1-160
( time timeout 300 python bench.py --impl reference ) > gpurun_out/z_bench_ref.log 2>&1
grep '^{' gpurun_out/z_bench_ref.log | cut -c1-160
export MPDO_BENCH_CUPROF=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/r2f_launches_bench.log 2>&1
unset MPDO_BENCH_CUPROF
wc -l gpurun_out/r2f_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chol_blocked|jacobi_cluster|jacobi_persistent_cols" -c 18 -f -o gpurun_out/r2f_fact python tools/ncu_targets_r2.py > gpurun_out/r2f_fact.log 2>&1
tail -2 gpurun_out/r2f_fact.log; ls -la gpurun_out/r2f_fact.ncu-rep
rm -f gpurun_out/z_configs.jsonl
( time timeout 900 python bench_configs.py --configs 1,2,3,4,5 --out gpurun_out/z_configs.jsonl ) > gpurun_out/z_configs.log 2>&1
cut -c1-260 gpurun_out/z_configs.jsonl; tail -4 gpurun_out/z_configs.log | cut -c1-200
