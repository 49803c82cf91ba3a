#!/bin/bash
# 8 GPUs: slab tests (multi-process, CUDA IPC) + the N = 4M strong-scaling point at 8, 4 GPUs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; tail -n 8 gpurun_out/r02h_pytest.log
for n in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --workload 4M --steps 100 --warmup 10 --e2e-steps 3 > gpurun_out/r02h_bench$n.json 2> gpurun_out/r02h_bench$n.err; tail -c 2500 gpurun_out/r02h_bench$n.json; tail -5 gpurun_out/r02h_bench$n.err
done
