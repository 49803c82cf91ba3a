#!/bin/bash
# one GPU: routing thresholds (cells_max_occupancy, cells_mid_occupancy) around the one-wave limit of the fused launch
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r02t_$tag.json 2> gpurun_out/r02t_$tag.err; python tools/benchrow.py $tag gpurun_out/r02t_$tag.json; }
for w in 512k 256k 1M; do
for mm in "2 3" "3 3" "2 15" "2 4" "3 4" "3 15" "1 15" "1 3" "4 4"; do
set -- $mm
run ${w}_m$1_$2 --workload $w --steps 100 --warmup 10 --cells-max-occupancy $1 --cells-mid-occupancy $2
done
done
