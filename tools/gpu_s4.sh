mkdir -p gpurun_out
( time timeout 800 python -m pytest tests -m gpu -x -q ) > gpurun_out/s4_pytest.log 2>&1
tail -3 gpurun_out/s4_pytest.log
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-cfg4 --no-unfused 2> gpurun_out/s4_bench.err | grep '^{"metric' > gpurun_out/s4_bench_cluster.json
MPDO_JACOBI_NOCLUSTER=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-cfg4 --no-unfused 2>> gpurun_out/s4_bench.err | grep '^{"metric' > gpurun_out/s4_bench_old.json
python - <<'PY'
import json
for f in ('cluster','old'):
    d=json.load(open('gpurun_out/s4_bench_%s.json'%f)); print(f, round(d['value'],1), d['ms_each_step'], 'e2e', round(d['e2e']['value'],1))
PY
