mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc_apply.py tests/test_gpu_prims.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/bench_tc_apply.py > gpurun_out/t14_tc.log 2>&1; tail -12 gpurun_out/t14_tc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tc_apply" -c 4 -o gpurun_out/r2_tc_after python tools/ncu_targets_r2.py > gpurun_out/t14_ncu.log 2>&1
tail -2 gpurun_out/t14_ncu.log
