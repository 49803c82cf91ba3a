#!/bin/bash
# 2 GPUs: wide mirror stores -- virtual slabs on one GPU, multi-process slab tests, pass time at 2 x 524288 (before: 56.7 us)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_slab.py -m gpu -x -q > gpurun_out/r03f_pytest.log 2>&1; tail -n 4 gpurun_out/r03f_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --workload 1M --steps 60 --warmup 8 --e2e-steps 2 > gpurun_out/r03f_1M.json 2> gpurun_out/r03f_1M.err; python tools/benchrow.py 2x512k gpurun_out/r03f_1M.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 2 --steps 40 --warmup 8 --e2e-steps 2 > gpurun_out/r03f_4M.json 2> gpurun_out/r03f_4M.err; python tools/benchrow.py 2gpu_4M gpurun_out/r03f_4M.json
