#!/bin/bash
# quick scan of the sweep-kernel tile width on two workloads (runs under gpurun)
for w in 64k 1M; do
  for l in 0 1 8 16; do
    steps=400; [ $w = 1M ] && steps=60
    python bench.py --workload $w --lanes $l --steps $steps --warmup 5 --no-cpu-baseline --e2e-steps 3 --relax-sweeps 200 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
k=d['kernels']
print('$w lanes=$l', 'value=%.3e warm=%.3e ms/step=%.4f sweep_pass_ms=%.4f cells_ms=%.4f pairchk/s=%.3e' % (d['value'], d['value_l2_warm'], d['ms_per_step'], k['sweep_pass']['ms_per_launch'], k['cell_build']['ms_per_build'], d['pair_checks_per_s_l2_warm']))"
  done
done
