mkdir -p gpurun_out
for v in notopk1 notopk2 default1 default2; do
case $v in notopk*) export MPDO_NO_TOPK=1;; *) unset MPDO_NO_TOPK;; esac
timeout 300 python bench.py --steps 20 --warmup 3 --no-cfg4 --no-unfused --no-cpu-baseline > gpurun_out/t10_bench_$v.json 2> gpurun_out/t10_bench_$v.err
python -c "
import json;d=json.load(open('gpurun_out/t10_bench_$v.json'));print('$v', round(d['value'],1),round(d['ms_per_step'],2),d['ms_each_step'],round(d['e2e']['value'],1), d['gpu_launches'])"
done
