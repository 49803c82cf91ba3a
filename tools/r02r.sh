#!/bin/bash
# one GPU: parity tests of the bond count + ncu lines of K3 at 4M / 32k
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; tail -n 12 gpurun_out/r02r_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:k_bonds --csv --log-file gpurun_out/r02r_aux_4M.csv python tools/aux_kernels.py 4194304 0.6 > gpurun_out/r02r_aux_4M.log 2>&1
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:k_bonds --csv --log-file gpurun_out/r02r_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02r_aux_32k.log 2>&1
python tools/launch_table.py gpurun_out/r02r_aux_4M.csv 4194304; python tools/launch_table.py gpurun_out/r02r_aux_32k.csv 32768
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bonds_screen -c 1 -f -o gpurun_out/r02r_screen_4M python tools/aux_kernels.py 4194304 0.6 0 > gpurun_out/r02r_ncu_full.log 2>&1
