mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"chol_blocked|chol_cluster|chol_small|jacobi" --csv --log-file gpurun_out/t17_chol_list.csv python tools/bench_chol2.py > /dev/null 2>&1
python - <<'P'
import csv,collections
rows=[l for l in open('gpurun_out/t17_chol_list.csv') if l.startswith('"')]
r=list(csv.DictReader(rows))
# group by kernel name + grid size
agg=collections.OrderedDict()
for x in r:
    k=(x['Kernel Name'].split('(')[0][-40:], x['Grid Size'], x['Block Size'])
    agg.setdefault(k,[]).append(float(x['Metric Value'].replace(',','')))
for k,v in agg.items(): print(k, len(v), 'min %.1f us'%(min(v)/1e3), 'median %.1f us'%(sorted(v)[len(v)//2]/1e3))
P
bash tools/gpu_t11.sh
