#!/bin/bash
# Runs on the GPU box (under gpurun): bench line, ncu launch list, one full capture of the sweep kernel.
# usage: tools/gpu_profile.sh <tag> [extra bench args]
tag=${1:-r01}; shift
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -c 3000 gpurun_out/bench_${tag}.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 40 --warmup 3 --no-cpu-baseline --e2e-steps 3 "$@" > gpurun_out/ncu_launches_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep_pass -s 80 -c 8 -f -o gpurun_out/sweep_${tag} \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 3 "$@" > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out | tail -8
