mkdir -p gpurun_out
timeout 300 python tools/prof_outliers.py 80 > gpurun_out/t12_outliers.log 2>&1
grep -c total gpurun_out/t12_outliers.log
awk '/total/ {print $3}' gpurun_out/t12_outliers.log | sort -n | tail -8 | tr '\n' ' '; echo
grep -B1 -A16 OUTLIER gpurun_out/t12_outliers.log | cut -c1-250 | head -80
grep "num_device_alloc" gpurun_out/t12_outliers.log | cut -c1-250 | head
