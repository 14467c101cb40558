mkdir -p gpurun_out
timeout 250 python tools/bench_jacobi.py > gpurun_out/s2_jac_cluster.log 2>&1
MPDO_JACOBI_NOCLUSTER=1 timeout 250 python tools/bench_jacobi.py > gpurun_out/s2_jac_old.log 2>&1
timeout 300 python -m pytest tests/test_gpu_prims.py -x -q -m gpu 2>&1 | tail -5 > gpurun_out/s2_pytest.log
cat gpurun_out/s2_jac_cluster.log | tail -30
