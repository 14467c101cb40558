mkdir -p gpurun_out
LAYER=12 timeout 200 python tools/prof_critical.py > gpurun_out/t5_critical12.log 2>&1
grep -v "qr_step\|bond_svd" gpurun_out/t5_critical12.log | cut -c1-220
