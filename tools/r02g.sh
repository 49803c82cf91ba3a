#!/bin/bash
# 2 GPUs: multi-process slab tests (CUDA IPC, device-side phase flags) + group over devices + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_group.py -m gpu -x -q -k "slab_matches or over_devices" > gpurun_out/r02g_pytest.log 2>&1; tail -n 15 gpurun_out/r02g_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 8 --e2e-steps 3 > gpurun_out/r02g_bench2.json 2> gpurun_out/r02g_bench2.err; tail -c 1800 gpurun_out/r02g_bench2.json; tail -5 gpurun_out/r02g_bench2.err
