#!/bin/bash
# one GPU: tile kernel against the fused launch (one warp per block) below the 150k routing threshold
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 2 "$@" > gpurun_out/r03r_$tag.json 2> gpurun_out/r03r_$tag.err; python tools/benchrow.py $tag gpurun_out/r03r_$tag.json; }
run 64k_tile --workload 64k --steps 200 --warmup 20
run 64k_m1_2 --workload 64k --steps 200 --warmup 20 --cells-max-occupancy 1 --cells-mid-occupancy 2
run 64k_m2_3 --workload 64k --steps 200 --warmup 20 --cells-max-occupancy 2 --cells-mid-occupancy 3
run 128k_tile --workload 128k --steps 200 --warmup 20
run 128k_m1_2 --workload 128k --steps 200 --warmup 20 --cells-max-occupancy 1 --cells-mid-occupancy 2
run 128k_m2_3 --workload 128k --steps 200 --warmup 20 --cells-max-occupancy 2 --cells-mid-occupancy 3
