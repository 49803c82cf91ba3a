#!/bin/bash
# one GPU: parity + group tests with the dataflow half-sweep launch, then sizes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; tail -n 15 gpurun_out/r02s_pytest.log
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r02s_$tag.json 2> gpurun_out/r02s_$tag.err; python tools/benchrow.py $tag gpurun_out/r02s_$tag.json; }
run 512k --workload 512k --steps 100 --warmup 10
run 512k_perpass --workload 512k --steps 100 --warmup 10 --cell-lanes -2
run 4M --workload 4M --steps 40 --warmup 5
run 1M --workload 1M --steps 60 --warmup 8
run 256k --workload 256k --steps 200 --warmup 10
run 64k_m1 --workload 64k --steps 400 --warmup 20 --cells-max-occupancy 1 --cells-mid-occupancy 2
run 64k_m2 --workload 64k --steps 400 --warmup 20 --cells-max-occupancy 2 --cells-mid-occupancy 3
run 64k --workload 64k --steps 400 --warmup 20
