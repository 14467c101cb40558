#!/bin/bash
mkdir -p gpurun_out
N=2
T0=$(date +%s)
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/y_bench_n$N.json 2> gpurun_out/y_bench_n$N.err
echo "N=$N bench wall $(( $(date +%s) - T0 )) s"; tail -3 gpurun_out/y_bench_n$N.err
python - <<P
import json
d=json.loads([l for l in open('gpurun_out/y_bench_n2.json') if l.startswith('{')][-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],2),'e2e',round(d['e2e']['value'],1))
print('cfg4',d['cfg4']['value'],d['cfg4']['e2e']['value'], d['cfg4'].get('circuits_total'))
P
