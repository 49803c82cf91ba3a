#!/bin/bash
# one GPU: GPU test suite, both bench arms as the driver runs them, ncu launch list of the 4M step (time + DRAM bytes, caches
# not flushed), auxiliary kernels at 4M and 32k, one full capture of the fused sweep kernel at 4M
mkdir -p gpurun_out
( time timeout 1300 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02m_pytest.log 2>&1; tail -n 8 gpurun_out/r02m_pytest.log
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02m_ref.json 2> gpurun_out/r02m_ref.err; tail -c 700 gpurun_out/r02m_ref.json; tail -4 gpurun_out/r02m_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; tail -c 6000 gpurun_out/r02m_bench.json; tail -6 gpurun_out/r02m_bench.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --cache-control none --clock-control none -s 200 -c 300 --csv --log-file gpurun_out/r02m_launches_4M.csv python bench.py --workload 4M --steps 8 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02m_ncu_launches.log 2>&1
timeout 600 ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02m_aux_4M.csv python tools/aux_kernels.py 4194304 0.6 > gpurun_out/r02m_aux_4M.log 2>&1
timeout 300 ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02m_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02m_aux_32k.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fused -s 40 -c 2 -f -o gpurun_out/r02m_fused_4M python bench.py --workload 4M --steps 6 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02m_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
