#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size
for w in 512k 4M; do
ncu --metrics $M --cache-control none --clock-control none -k regex:k_sweep_cells -s 48 -c 48 --csv --log-file gpurun_out/r02l_split_$w.csv python bench.py --workload $w --cell-lanes -1 --steps 6 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02l_$w.log 2>&1
done
python bench.py --workload 512k --cell-lanes -1 --steps 100 --warmup 10 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02l_512k_split.json 2>gpurun_out/r02l_512k_split.err; python tools/benchrow.py 512k_split gpurun_out/r02l_512k_split.json
