#!/bin/bash
# Final evidence of round 2 (session 4): ncu launch list of the timed region of bench.py on the final code, and full
# captures of the factorisation kernels incl. the blocked Cholesky.
mkdir -p gpurun_out
export MPDO_BENCH_CUPROF=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/r2f_launches_bench.log 2>&1
unset MPDO_BENCH_CUPROF
wc -l gpurun_out/r2f_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"chol_blocked|jacobi_cluster" -c 12 -f -o gpurun_out/r2f_fact python tools/ncu_targets_r2.py > gpurun_out/r2f_fact.log 2>&1
tail -2 gpurun_out/r2f_fact.log; ls -la gpurun_out/r2f_fact.ncu-rep
