mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/t1_pytest.log
cat gpurun_out/t1_pytest.log
timeout 200 python tools/prof_host.py > gpurun_out/t1_host.log 2>&1
cat gpurun_out/t1_host.log
timeout 300 python bench.py --steps 8 --warmup 3 > gpurun_out/t1_bench.json 2> gpurun_out/t1_bench.err
tail -c 600 gpurun_out/t1_bench.err
python -c "
import json;d=json.load(open('gpurun_out/t1_bench.json'));print(d['value'],d['ms_per_step'],d['ms_each_step'],d['e2e'],d.get('cfg4'))"
