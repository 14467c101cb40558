mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_native_engine.py -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_t11.sh
