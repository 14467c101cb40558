mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_env_sweep.py -x -q -m gpu -s 2>&1 | tail -25 > gpurun_out/t3_env.log
cat gpurun_out/t3_env.log
for e in 0 1; do
MPDO_ENV_SWEEP=$e timeout 300 python bench.py --steps 8 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/t3_bench_e$e.json 2> gpurun_out/t3_bench_e$e.err
tail -c 400 gpurun_out/t3_bench_e$e.err
python -c "
import json;d=json.load(open('gpurun_out/t3_bench_e$e.json'));print('env $e', d['value'],d['ms_per_step'],d['ms_each_step'],d['e2e']['value'], d['gpu_launches'])"
done
MPDO_ENV_SWEEP=1 timeout 200 python tools/prof_host.py > gpurun_out/t3_host.log 2>&1
tail -4 gpurun_out/t3_host.log
