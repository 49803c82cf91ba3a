#!/bin/bash
# one GPU: K3 at small sizes after the latency fixes, compute-sanitizer memcheck / racecheck / synccheck over every kernel class
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_sizes.py -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1; tail -n 5 gpurun_out/r02u_pytest.log
M=gpu__time_duration.sum
timeout 300 ncu --metrics $M --cache-control none --clock-control none -k regex:k_bonds --csv --log-file gpurun_out/r02u_aux_32k.csv python tools/aux_kernels.py 32768 0.3 > gpurun_out/r02u_aux_32k.log 2>&1
python tools/launch_table.py gpurun_out/r02u_aux_32k.csv
timeout 300 python bench.py --workload npt32k --steps 300 --warmup 20 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02u_npt32k.json 2> gpurun_out/r02u_npt32k.err; python tools/benchrow.py npt32k gpurun_out/r02u_npt32k.json
for tool in memcheck racecheck synccheck; do
( time timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_run.py ) > gpurun_out/r02u_sanitizer_$tool.log 2>&1; tail -n 6 gpurun_out/r02u_sanitizer_$tool.log
done
