#!/bin/bash
# single-GPU throughput with the fused pass kernel: sizes, routing thresholds
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py --no-cpu-baseline --e2e-steps 3 "$@" > gpurun_out/r02e_$tag.json 2> gpurun_out/r02e_$tag.err; python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/r02e_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    k=d["kernels"]
    print(f"{sys.argv[1]:28s} value={d['value']:.4e} ms/step={d['ms_per_step']:.4f} pass_ms={k['sweep_pass']['ms_per_launch']:.4f} cells_ms={k['cell_build']['ms_per_build']:.4f} rng_ms={k['proposals']['ms_per_launch']:.4f} bonds_ms={k['bond_count']['ms']/max(k['bond_count']['launches']//2,1):.4f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(f"gpurun_out/r02e_{sys.argv[1]}.err").read()[-400:])
PY
}
run 64k --workload 64k --steps 400 --warmup 20
run 64k_fused_m1 --workload 64k --steps 400 --warmup 20 --cells-max-occupancy 1 --cells-mid-occupancy 2
run 64k_fused_m0_4 --workload 64k --steps 400 --warmup 20 --cells-max-occupancy 1 --cells-mid-occupancy 1
run 256k --workload 256k --steps 200 --warmup 10
run 512k --workload 512k --steps 100 --warmup 10
run 512k_m1_2 --workload 512k --steps 100 --warmup 10 --cells-max-occupancy 1 --cells-mid-occupancy 2
run 512k_m1_3 --workload 512k --steps 100 --warmup 10 --cells-max-occupancy 1 --cells-mid-occupancy 3
run 512k_m2_2 --workload 512k --steps 100 --warmup 10 --cells-max-occupancy 2 --cells-mid-occupancy 2
run 512k_m2_4 --workload 512k --steps 100 --warmup 10 --cells-max-occupancy 2 --cells-mid-occupancy 4
run 512k_m3_3 --workload 512k --steps 100 --warmup 10 --cells-max-occupancy 3 --cells-mid-occupancy 3
run 1M --workload 1M --steps 60 --warmup 8
run 4M --workload 4M --steps 30 --warmup 5
