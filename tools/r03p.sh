#!/bin/bash
# one GPU: the cells kernels with one warp per block (12 blocks per SM instead of 6 of two warps)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r03p_pytest.log 2>&1; tail -n 3 gpurun_out/r03p_pytest.log
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 2 "$@" > gpurun_out/r03p_$tag.json 2> gpurun_out/r03p_$tag.err; python tools/benchrow.py $tag gpurun_out/r03p_$tag.json; }
for w in 256k 512k 576k 1M 4M; do run $w --workload $w --steps 60 --warmup 10; done
