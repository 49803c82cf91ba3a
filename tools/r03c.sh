#!/bin/bash
# 2 GPUs, timing experiment: the pass with and without the mirror stores into the neighbour's halo (the second run's physics is wrong)
mkdir -p gpurun_out
for tag in mirror nomirror; do
if [ $tag = nomirror ]; then export PP_DEBUG_NO_MIRROR=1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload 1M --steps 60 --warmup 8 --e2e-steps 2 > gpurun_out/r03c_$tag.json 2> gpurun_out/r03c_$tag.err; python tools/benchrow.py $tag gpurun_out/r03c_$tag.json
done
unset PP_DEBUG_NO_MIRROR
timeout 300 python bench.py --workload 512k --steps 100 --warmup 10 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r03c_512k.json 2> gpurun_out/r03c_512k.err; python tools/benchrow.py 512k_1gpu gpurun_out/r03c_512k.json
