#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r03s_pytest.log 2>&1; tail -n 6 gpurun_out/r03s_pytest.log
timeout 200 python bench.py --workload 128k --steps 200 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 2 > gpurun_out/r03s_128k.json 2> gpurun_out/r03s_128k.err; python tools/benchrow.py 128k_auto gpurun_out/r03s_128k.json
