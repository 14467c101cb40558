mkdir -p gpurun_out
MPDO_ENV_SWEEP=1 timeout 200 python tools/prof_host.py > gpurun_out/t4_host.log 2>&1
tail -3 gpurun_out/t4_host.log | cut -c1-1500
