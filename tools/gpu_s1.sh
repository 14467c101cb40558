mkdir -p gpurun_out
timeout 200 python tools/prof_critical.py > gpurun_out/s1_critical.log 2>&1
timeout 200 python tools/prof_sweep_kernels.py > gpurun_out/s1_phase_kernels.log 2>&1
timeout 200 python tools/bench_chol2.py > gpurun_out/s1_chol2.log 2>&1
tail -3 gpurun_out/s1_chol2.log
