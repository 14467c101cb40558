mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 200 python tools/bench_chol2.py > gpurun_out/t19_chol.log 2>&1; cat gpurun_out/t19_chol.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/t19_bench.json 2> gpurun_out/t19_bench.err
python -c "
import json;d=json.load(open('gpurun_out/t19_bench.json'));print(round(d['value'],1),round(d['ms_per_step'],2),d['ms_each_step'],round(d['e2e']['value'],1), d['gpu_launches'])"
