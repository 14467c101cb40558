#!/bin/bash
mkdir -p gpurun_out
MPDO_NO_DMMA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 2 -c 2 -f -o gpurun_out/c4_contract_scalar python tools/ncu_contract.py > gpurun_out/c4_ncu_scalar.log 2>&1
MPDO_DMMA_ALL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:contract_kernel -s 2 -c 2 -f -o gpurun_out/c4_contract_dmma python tools/ncu_contract.py > gpurun_out/c4_ncu_dmma.log 2>&1
ls -la gpurun_out/*.ncu-rep
