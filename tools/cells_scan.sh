#!/bin/bash
# scan of the thread-per-cell kernel's lanes per cell and routing threshold (runs under gpurun)
# usage: tools/cells_scan.sh "<workload> <lanes> <cells-max-occupancy> <cell-lanes>" ...
for cfg in "$@"; do
  set -- $cfg
  steps=300; [ $1 = 1M ] && steps=60; [ $1 = 512k ] && steps=100
  python bench.py --workload $1 --lanes $2 --cells-max-occupancy $3 --cell-lanes $4 --steps $steps --warmup 5 --no-cpu-baseline --e2e-steps 3 --relax-sweeps 200 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['kernels']
print('$1 lanes=$2 M=$3 cell_lanes=$4', 'value=%.3e ms/step=%.4f sweep_pass_ms=%.4f' % (d['value'], d['ms_per_step'], k['sweep_pass']['ms_per_launch']))"
done
