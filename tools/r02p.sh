#!/bin/bash
# one GPU: parity tests, K3 (screen + exact kernels) ncu lines at 4M / 32k, sweep sizes with the q-side FP32 screen and edge-cell sub-bins
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r02p_pytest.log 2>&1; tail -n 12 gpurun_out/r02p_pytest.log
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --metrics $M --cache-control none --clock-control none -k regex:k_bonds --csv --log-file gpurun_out/r02p_aux_4M.csv python tools/aux_kernels.py 4194304 0.6 > gpurun_out/r02p_aux_4M.log 2>&1
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:k_bonds --csv --log-file gpurun_out/r02p_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02p_aux_32k.log 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/r02p_aux_4M.csv", "gpurun_out/r02p_aux_32k.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    for r in rows[1:13]:
        print(f[-10:], r[4][:18], r[-3], r[-1])
PY
run() { tag=$1; shift; timeout 600 python bench.py --no-cpu-baseline --no-configs --e2e-steps 3 "$@" > gpurun_out/r02p_$tag.json 2> gpurun_out/r02p_$tag.err; python tools/benchrow.py $tag gpurun_out/r02p_$tag.json; }
run 512k --workload 512k --steps 100 --warmup 10
run 4M --workload 4M --steps 40 --warmup 5
run 1M --workload 1M --steps 60 --warmup 8
run 256k --workload 256k --steps 200 --warmup 10
