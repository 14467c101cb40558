mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_prims.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/s6_pytest.log
cat gpurun_out/s6_pytest.log | tail -5
timeout 200 python tools/bench_chol2.py > gpurun_out/s6_chol_blocked.log 2>&1
MPDO_CHOL_NOBLOCK=1 timeout 200 python tools/bench_chol2.py > gpurun_out/s6_chol_old.log 2>&1
paste gpurun_out/s6_chol_blocked.log gpurun_out/s6_chol_old.log | cut -c1-200
