#!/bin/bash
# one GPU: K3 at small sizes with full-width blocks; NPT / Lx rates; parity of the counts
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r03e_pytest.log 2>&1; tail -n 3 gpurun_out/r03e_pytest.log
M=gpu__time_duration.sum
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:k_bonds --csv --log-file gpurun_out/r03e_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r03e_aux_32k.log 2>&1
python tools/launch_table.py gpurun_out/r03e_aux_32k.csv
run() { tag=$1; shift; timeout 300 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r03e_$tag.json 2> gpurun_out/r03e_$tag.err; python tools/benchrow.py $tag gpurun_out/r03e_$tag.json; }
run npt32k --workload npt32k --steps 300 --warmup 20
run lx32k --workload lx32k --steps 300 --warmup 20
