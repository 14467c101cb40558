mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t6_pytest.log
cat gpurun_out/t6_pytest.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/t6_bench.json 2> gpurun_out/t6_bench.err
tail -c 400 gpurun_out/t6_bench.err
python -c "
import json;d=json.load(open('gpurun_out/t6_bench.json'));print(d['value'],d['ms_per_step'],d['ms_each_step'],d['e2e']['value'], d['gpu_launches'])"
MPDO_GROUPING=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/t6_bench_g1.json 2> gpurun_out/t6_bench_g1.err
python -c "
import json;d=json.load(open('gpurun_out/t6_bench_g1.json'));print('grouping', d['value'],d['ms_per_step'],d['ms_each_step'],d['e2e']['value'], d['gpu_launches'])"
timeout 200 python tools/prof_host.py > gpurun_out/t6_host.log 2>&1
tail -5 gpurun_out/t6_host.log | cut -c1-400
