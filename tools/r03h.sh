#!/bin/bash
# one GPU: the colour pass around the size where the working set of the sweep (proposals + FP64 state + shadows) meets the 126 MB of L2
mkdir -p gpurun_out
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 2 "$@" > gpurun_out/r03h_$tag.json 2> gpurun_out/r03h_$tag.err; python tools/benchrow.py $tag gpurun_out/r03h_$tag.json; }
for w in 448k 512k 576k 640k 768k 1M; do run $w --workload $w --steps 100 --warmup 10; done
