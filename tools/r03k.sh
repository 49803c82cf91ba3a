#!/bin/bash
# one GPU: routing thresholds at N = 589824 (the size a slab engine of the 8-GPU run holds), where the pass has jumped from 46 to 56 us
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 2 "$@" > gpurun_out/r03k_$tag.json 2> gpurun_out/r03k_$tag.err; python tools/benchrow.py $tag gpurun_out/r03k_$tag.json; }
for mm in "2 3" "2 2" "2 4" "2 15" "3 3" "3 4" "1 2" "1 3"; do
set -- $mm
run 576k_m$1_$2 --workload 576k --steps 60 --warmup 10 --cells-max-occupancy $1 --cells-mid-occupancy $2
done
