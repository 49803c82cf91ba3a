#!/bin/bash
# one GPU: proposal buffer as streaming lines in L2 (access policy window): the cliff between 512k and 576k, with and without
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r03i_pytest.log 2>&1; tail -n 3 gpurun_out/r03i_pytest.log
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 2 "$@" > gpurun_out/r03i_$tag.json 2> gpurun_out/r03i_$tag.err; python tools/benchrow.py $tag gpurun_out/r03i_$tag.json; }
for w in 512k 576k 640k 768k 1M 4M; do run $w --workload $w --steps 60 --warmup 10; done
export PP_NO_STREAM_POLICY=1
for w in 576k 1M; do run ${w}_nopolicy --workload $w --steps 60 --warmup 10; done
