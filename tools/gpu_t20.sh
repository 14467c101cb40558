mkdir -p gpurun_out
timeout 200 python tools/prof_host.py > gpurun_out/t20_host.log 2>&1
tail -8 gpurun_out/t20_host.log | cut -c1-600
LAYER=12 timeout 200 python tools/prof_critical.py > gpurun_out/t20_critical12.log 2>&1
grep -v "bond_svd\|qr_step" gpurun_out/t20_critical12.log | cut -c1-200 | head -30
