#!/bin/bash
# 8 GPUs: the N = 4M strong-scaling point (driver's launch line)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 100 --warmup 10 --e2e-steps 3 > gpurun_out/r02q_bench8.json 2> gpurun_out/r02q_bench8.err; tail -c 3500 gpurun_out/r02q_bench8.json; tail -5 gpurun_out/r02q_bench8.err
