#!/bin/bash
# one GPU: where the FCC crystal (configs[4]) spends its pass; round-trip e2e with the pinned download
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_sweep_fused -s 24 -c 1 -f -o gpurun_out/r02x_fused_fcc python bench.py --workload fcc256k --steps 6 --warmup 3 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02x_ncu.log 2>&1
timeout 300 python bench.py --workload fcc256k --steps 100 --warmup 10 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02x_fcc.json 2> gpurun_out/r02x_fcc.err; python tools/benchrow.py fcc gpurun_out/r02x_fcc.json
timeout 300 python bench.py --workload 4M --steps 20 --warmup 5 --no-cpu-baseline --no-configs --e2e-steps 5 > gpurun_out/r02x_4M.json 2> gpurun_out/r02x_4M.err; python tools/benchrow.py 4M gpurun_out/r02x_4M.json; python -c "
import json; d=json.loads(open('gpurun_out/r02x_4M.json').read().strip().splitlines()[-1]); print('roundtrip', d['e2e_roundtrip'])"
