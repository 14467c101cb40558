mkdir -p gpurun_out
export MPDO_BENCH_CUPROF=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/r2_launches_bench.log 2>&1
unset MPDO_BENCH_CUPROF
wc -l gpurun_out/r2_launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_apply|chol_small|chol_cluster|jacobi_cluster|jacobi_kernel|contract_kernel" -c 40 -o gpurun_out/r2_targets python tools/ncu_targets_r2.py > gpurun_out/r2_targets.log 2>&1
tail -3 gpurun_out/r2_targets.log; ls -la gpurun_out/r2_targets.ncu-rep
