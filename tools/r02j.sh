#!/bin/bash
# one GPU: GPU test suite with the box-fraction shadows + sizes scan of the rewritten thread-per-cell kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02j_pytest.log 2>&1; tail -n 6 gpurun_out/r02j_pytest.log
run() { tag=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r02j_$tag.json 2> gpurun_out/r02j_$tag.err; python tools/benchrow.py $tag gpurun_out/r02j_$tag.json; }
run 512k --workload 512k --steps 100 --warmup 10
run 512k_m3 --workload 512k --steps 100 --warmup 10 --cells-max-occupancy 3 --cells-mid-occupancy 4
run 512k_m1 --workload 512k --steps 100 --warmup 10 --cells-max-occupancy 1 --cells-mid-occupancy 3
run 4M --workload 4M --steps 40 --warmup 5
run 1M --workload 1M --steps 60 --warmup 8
run 256k --workload 256k --steps 200 --warmup 10
run 64k --workload 64k --steps 400 --warmup 20
run 64k_m1 --workload 64k --steps 400 --warmup 20 --cells-max-occupancy 1 --cells-mid-occupancy 2
run 64k_m2 --workload 64k --steps 400 --warmup 20 --cells-max-occupancy 2 --cells-mid-occupancy 3
run fcc --workload fcc256k --steps 100 --warmup 10
