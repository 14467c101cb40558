mkdir -p gpurun_out
for g in 0 1; do
MPDO_GROUPING=$g timeout 300 python bench.py --steps 8 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/t2_bench_g$g.json 2> gpurun_out/t2_bench_g$g.err
python -c "
import json;d=json.load(open('gpurun_out/t2_bench_g$g.json'));print('grouping $g', d['value'],d['ms_per_step'],d['ms_each_step'],d['e2e']['value'], d['gpu_launches'])"
done
timeout 400 python tools/prof_cfg4.py > gpurun_out/t2_cfg4.log 2>&1
cut -c1-200 gpurun_out/t2_cfg4.log | head -60
