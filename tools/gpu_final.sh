#!/bin/bash
# Round-end evidence run: GPU tests, both bench arms, the ncu launch list of the bench command, full captures of the
# heavy kernels, and the five configurations.
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/f_pytest.log 2>&1
grep -E "passed|failed|error" gpurun_out/f_pytest.log | tail -2
( time timeout 400 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/f_bench_ref.log 2>&1
( time timeout 400 python bench.py ) > gpurun_out/f_bench.log 2>&1
grep '^{' gpurun_out/f_bench.log | cut -c1-220
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f_launches_full.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/f_ncu_bench.log 2>&1
wc -l gpurun_out/f_launches_full.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"chol_kernel|jacobi_persistent_reg|contract_kernel" -c 8 -f -o gpurun_out/f_kernels python tools/ncu_targets.py > gpurun_out/f_ncu_kernels.log 2>&1
rm -f gpurun_out/f_configs.jsonl
( time timeout 900 python bench_configs.py --configs 1,2,4 --out gpurun_out/f_configs.jsonl ) > gpurun_out/f_configs.log 2>&1
cut -c1-200 gpurun_out/f_configs.jsonl
