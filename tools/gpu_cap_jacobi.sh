mkdir -p gpurun_out
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"jacobi_persistent" -s 1 -c 1 -f -o gpurun_out/j_jacobi python tools/ncu_targets.py > gpurun_out/j_ncu.log 2>&1
ls -la gpurun_out/j_jacobi.ncu-rep
