#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; tail -n 3 gpurun_out/r02c_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02c_aux_4M.csv python tools/aux_kernels.py 4194304 0.6 > gpurun_out/r02c_aux_4M.log 2>&1
ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02c_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02c_aux_32k.log 2>&1
grep -h "k_total_bonds" gpurun_out/r02c_aux_4M.csv | head -4; grep -h "k_total_bonds" gpurun_out/r02c_aux_32k.csv | head -4
