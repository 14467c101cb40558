#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_prims.py -m gpu -x -q 2>&1 | tail -2
echo "== cluster"; timeout 120 python tools/bench_chol.py 2>&1 | tee gpurun_out/c20_chol_cluster.log | grep -v "^\[chol"
echo "== global barrier"; MPDO_CHOL_NOCLUSTER=1 timeout 120 python tools/bench_chol.py 2>&1 | tee gpurun_out/c20_chol_global.log | grep -v "^\[chol"
export MPDO_CHOL_PROFILE=1
echo "== cluster prof"; timeout 120 python tools/bench_chol.py 2>&1 | grep -E "^\[chol" | awk '{k=$2" "$3" "$4" "$5; if (!(k in seen)) {seen[k]=1; print}}'
echo "== global prof"; MPDO_CHOL_NOCLUSTER=1 timeout 120 python tools/bench_chol.py 2>&1 | grep -E "^\[chol" | awk '{k=$2" "$3" "$4" "$5; if (!(k in seen)) {seen[k]=1; print}}'
