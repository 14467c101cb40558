mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/t11_bench.json 2> gpurun_out/t11_bench.err
echo "full bench wall $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/t11_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/t11_bench.json'))
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1),'unfused',d.get('value_one_split_per_gate'))
print('steps',d['ms_each_step'])
print('cfg4',d['cfg4']['value'],d['cfg4']['e2e']['value'])
print('cpu',d['cpu_baseline'])
r=d['roofline']; print('roofline frac',r['frac'],'share',r['share_of_step_device_time']); print(json.dumps(d['factorisation_kernels'])[:1500])
P
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/t11_ref.json 2> gpurun_out/t11_ref.err
tail -2 gpurun_out/t11_ref.err; cut -c1-600 gpurun_out/t11_ref.json
