#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02f_aux_4M.csv python tools/aux_kernels.py 4194304 0.6 > gpurun_out/r02f_aux_4M.log 2>&1
ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02f_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02f_aux_32k.log 2>&1
grep -h "k_total_bonds" gpurun_out/r02f_aux_4M.csv | grep duration | head -2; grep -h "k_total_bonds" gpurun_out/r02f_aux_32k.csv | grep duration | head -2
ncu --set full --clock-control none --import-source on -k regex:k_sweep_fused -s 40 -c 2 -f -o gpurun_out/r02f_fused_512k python bench.py --workload 512k --steps 6 --warmup 3 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02f_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep
