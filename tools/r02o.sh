#!/bin/bash
# one GPU: parity tests of the bond count (pencil kernel) + its ncu line at 4M and 32k
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r02o_pytest.log 2>&1; tail -n 12 gpurun_out/r02o_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:k_total_bonds --csv --log-file gpurun_out/r02o_aux_4M.csv python tools/aux_kernels.py 4194304 0.6 > gpurun_out/r02o_aux_4M.log 2>&1
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:k_total_bonds --csv --log-file gpurun_out/r02o_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02o_aux_32k.log 2>&1
grep -h "k_total_bonds" gpurun_out/r02o_aux_4M.csv | grep duration | head -3; grep -h "k_total_bonds" gpurun_out/r02o_aux_4M.csv | grep dram | head -4; grep -h "k_total_bonds" gpurun_out/r02o_aux_32k.csv | grep duration | head -3
cat gpurun_out/r02o_aux_4M.log | tail -5
