#!/bin/bash
# one GPU: new tests (neighbour queries, K3 queue overflow), the whole GPU suite, both bench arms as the driver runs them
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "neighbour or overflow or gc_" > gpurun_out/r02w_new.log 2>&1; tail -n 12 gpurun_out/r02w_new.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02w_pytest.log 2>&1; tail -n 6 gpurun_out/r02w_pytest.log
( time timeout 600 python bench.py --impl reference ) > gpurun_out/r02w_ref.json 2> gpurun_out/r02w_ref.err; tail -c 400 gpurun_out/r02w_ref.json; tail -4 gpurun_out/r02w_ref.err
( time timeout 900 python bench.py ) > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err; python tools/benchrow.py default gpurun_out/r02w_bench.json; tail -6 gpurun_out/r02w_bench.err
