#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 2 -c 2 -f -o gpurun_out/c15_contract python tools/ncu_contract.py > gpurun_out/c15_ncu.log 2>&1
ls -la gpurun_out/c15*.ncu-rep
