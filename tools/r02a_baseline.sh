#!/bin/bash
# round-2 baseline on one B200: GPU test suite, 4M on one GPU, ncu launch lists (time + DRAM bytes, caches not flushed) of the auxiliary kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; tail -3 gpurun_out/r02a_pytest.log
python bench.py --workload 4M --steps 30 --warmup 5 --no-cpu-baseline --e2e-steps 3 > gpurun_out/r02a_bench_4M_1gpu.json 2> gpurun_out/r02a_bench_4M.err; tail -c 1500 gpurun_out/r02a_bench_4M_1gpu.json
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02a_aux_4M.csv python tools/aux_kernels.py 4194304 0.6 > gpurun_out/r02a_aux_4M.log 2>&1
ncu --metrics $M --cache-control none --clock-control none --csv --log-file gpurun_out/r02a_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02a_aux_32k.log 2>&1
tail -3 gpurun_out/r02a_aux_4M.log gpurun_out/r02a_aux_32k.log
