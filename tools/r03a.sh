#!/bin/bash
# one GPU: coalesced proposal records -- parity, then rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r03a_pytest.log 2>&1; tail -n 4 gpurun_out/r03a_pytest.log
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r03a_$tag.json 2> gpurun_out/r03a_$tag.err; python tools/benchrow.py $tag gpurun_out/r03a_$tag.json; }
run 4M --workload 4M --steps 30 --warmup 5
run 512k --workload 512k --steps 100 --warmup 10
run 64k --workload 64k --steps 400 --warmup 20
