#!/bin/bash
# one GPU: sweep parity (fluid + crystal, every configuration) with the group-cooperative exact path, then FCC / 512k / 4M rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r02y_pytest.log 2>&1; tail -n 6 gpurun_out/r02y_pytest.log
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r02y_$tag.json 2> gpurun_out/r02y_$tag.err; python tools/benchrow.py $tag gpurun_out/r02y_$tag.json; }
run fcc --workload fcc256k --steps 100 --warmup 10
run fcc_m0 --workload fcc256k --steps 100 --warmup 10 --cells-max-occupancy 0
run 512k --workload 512k --steps 100 --warmup 10
run 4M --workload 4M --steps 30 --warmup 5
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fused -s 24 -c 1 -f -o gpurun_out/r02y_fused_fcc python bench.py --workload fcc256k --steps 6 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02y_ncu.log 2>&1
