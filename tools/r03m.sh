#!/bin/bash
# 2 GPUs: balanced slabs (x_layers_multiple) -- virtual slabs on one GPU, multi-process slab tests, the 4M bench on two engines
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/r03m_pytest.log 2>&1; tail -n 4 gpurun_out/r03m_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 bench.py --gpus 2 --steps 40 --warmup 8 --e2e-steps 2 > gpurun_out/r03m_4M.json 2> gpurun_out/r03m_4M.err; python tools/benchrow.py 2gpu_4M gpurun_out/r03m_4M.json; tail -3 gpurun_out/r03m_4M.err
