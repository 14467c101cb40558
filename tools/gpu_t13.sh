mkdir -p gpurun_out
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/t13_bench_n2.json 2> gpurun_out/t13_bench_n2.err
echo "N=2 bench wall $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/t13_bench_n2.err
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/t13_bench_n2.json') if l.startswith('{')][-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))
print('steps',d['ms_each_step'])
print('cfg4',d['cfg4']['value'],d['cfg4']['e2e']['value'], d['cfg4'].get('circuits_total'))
P
