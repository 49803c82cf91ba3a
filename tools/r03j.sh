#!/bin/bash
# one GPU: DRAM bytes and time of the fused pass at 512k and 576k (caches not flushed by the profiler)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum
for w in 512k 576k; do
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:k_sweep_fused -s 48 -c 16 --csv --log-file gpurun_out/r03j_$w.csv python bench.py --workload $w --steps 8 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 2 > gpurun_out/r03j_$w.log 2>&1
python - $w <<'PY'
import csv, sys, collections
rows=[r for r in csv.reader(open(f"gpurun_out/r03j_{sys.argv[1]}.csv")) if len(r)>10]
h=rows[0]; mi,vi=h.index("Metric Name"),h.index("Metric Value")
agg=collections.defaultdict(list)
for r in rows[1:]: agg[r[mi]].append(float(r[vi].replace(",","")))
print(sys.argv[1], {k: round(sum(v)/len(v),1) for k,v in agg.items()})
PY
done
