#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 ) > gpurun_out/c16_bench_n2.log 2>&1
grep '^{' gpurun_out/c16_bench_n2.log | cut -c1-400
tail -3 gpurun_out/c16_bench_n2.log | cut -c1-300
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 ) > gpurun_out/c16_ref_n2.log 2>&1
grep '^{' gpurun_out/c16_ref_n2.log | cut -c1-300
timeout 300 python -m pytest tests/test_gpu_prims.py -m gpu -x -q 2>&1 | tail -2
