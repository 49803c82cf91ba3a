#!/bin/bash
# one GPU: Lx statistics test, sizes scan of the sweep, full capture of the fused kernel at 512k (the per-GPU share of the 8-GPU run)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_host_and_statistics.py -m gpu -x -q -s -k "lx_statistics" > gpurun_out/r02n_pytest.log 2>&1; tail -n 8 gpurun_out/r02n_pytest.log
run() { tag=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r02n_$tag.json 2> gpurun_out/r02n_$tag.err; python tools/benchrow.py $tag gpurun_out/r02n_$tag.json; }
run 512k --workload 512k --steps 100 --warmup 10
run 1M --workload 1M --steps 60 --warmup 8
run 256k --workload 256k --steps 200 --warmup 10
run 64k --workload 64k --steps 400 --warmup 20
run 64k_cells --workload 64k --steps 400 --warmup 20 --cells-max-occupancy 1 --cells-mid-occupancy 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fused -s 40 -c 2 -f -o gpurun_out/r02n_fused_512k python bench.py --workload 512k --steps 6 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02n_ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
