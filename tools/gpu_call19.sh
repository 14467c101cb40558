#!/bin/bash
mkdir -p gpurun_out
export MPDO_CHOL_PROFILE=1
echo "== cluster"; timeout 120 python tools/bench_chol.py 2>&1 | grep -E "^\[chol" | sort | uniq -c | sort -rn | head -14
echo "== global barrier"; MPDO_CHOL_NOCLUSTER=1 timeout 120 python tools/bench_chol.py 2>&1 | grep -E "^\[chol" | sort | uniq -c | sort -rn | head -14
