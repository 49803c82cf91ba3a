#!/bin/bash
# one GPU: the default bench line (4M headline + the other BASELINE configurations), reference arm first as the driver runs it
mkdir -p gpurun_out
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02i_ref.json 2> gpurun_out/r02i_ref.err; tail -c 600 gpurun_out/r02i_ref.json; tail -4 gpurun_out/r02i_ref.err
( time timeout 1500 python bench.py ) > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -c 1500 gpurun_out/r02i_bench.json; tail -12 gpurun_out/r02i_bench.err
