#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_sweep_fused -s 40 -c 2 -f -o gpurun_out/r02k_fused_512k python bench.py --workload 512k --steps 6 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02k_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
