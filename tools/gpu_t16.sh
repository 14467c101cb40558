mkdir -p gpurun_out
for i in 1 2 3; do
timeout 300 python bench.py --steps 20 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/t16_bench_$i.json 2> gpurun_out/t16_bench_$i.err
python -c "
import json;d=json.load(open('gpurun_out/t16_bench_$i.json'));print('run $i', round(d['value'],1),round(d['ms_per_step'],2),d['ms_each_step'],round(d['e2e']['value'],1), d['gpu_launches'])"
done
