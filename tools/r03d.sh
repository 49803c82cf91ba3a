#!/bin/bash
# one GPU: full capture of the K3 screen at N = 32768 (the NPT configuration)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_bonds_screen -c 2 -f -o gpurun_out/r03d_screen_32k python tools/aux_kernels.py 32768 0.3 1 > gpurun_out/r03d_ncu.log 2>&1
ls -la gpurun_out/r03d_screen_32k.ncu-rep
