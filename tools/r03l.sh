#!/bin/bash
mkdir -p gpurun_out
export PP_DEBUG_RANGES=1
for w in 448k 512k 576k; do
timeout 300 python bench.py --workload $w --steps 4 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 2 > gpurun_out/r03l_$w.json 2> gpurun_out/r03l_$w.err; grep "ranges" gpurun_out/r03l_$w.err | head -3
done
