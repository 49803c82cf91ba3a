#!/bin/bash
# 2 GPUs: where the sharded e2e step spends its time (PP_SLAB_TIMING), slab tests, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_slab.py tests/test_gpu_group.py -m gpu -x -q -k "slab_matches or over_devices" > gpurun_out/r02z_pytest.log 2>&1; tail -n 4 gpurun_out/r02z_pytest.log
PP_SLAB_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 40 --warmup 8 --e2e-steps 4 > gpurun_out/r02z_bench2.json 2> gpurun_out/r02z_bench2.err; python tools/benchrow.py 2gpu gpurun_out/r02z_bench2.json; grep "slab upload" gpurun_out/r02z_bench2.err | tail -8
