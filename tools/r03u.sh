#!/bin/bash
# one GPU: launch list of the 4M step at the end of the round (time + DRAM bytes, caches not flushed by the profiler)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 150 ncu --metrics $M --cache-control none --clock-control none -s 200 -c 220 --csv --log-file gpurun_out/r03u_launches_4M.csv python bench.py --workload 4M --steps 8 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 2 > gpurun_out/r03u_ncu.log 2>&1
python tools/launch_table.py gpurun_out/r03u_launches_4M.csv | head -16
