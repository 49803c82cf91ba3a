#!/bin/bash
# one GPU: GC parity + statistics tests, compute-sanitizer memcheck / racecheck / synccheck over every kernel class (tools/sanitize_run.py),
# then the whole GPU suite
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_host_and_statistics.py -m gpu -x -q -s -k "gc_ or sus_" > gpurun_out/r02v_gc.log 2>&1; tail -n 12 gpurun_out/r02v_gc.log
for tool in memcheck racecheck synccheck; do
( time timeout 900 compute-sanitizer --tool $tool --print-limit 400 python tools/sanitize_run.py ) > gpurun_out/r02v_sanitizer_$tool.log 2>&1; tail -n 5 gpurun_out/r02v_sanitizer_$tool.log
done
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/r02v_pytest.log 2>&1; tail -n 6 gpurun_out/r02v_pytest.log
timeout 300 python bench.py --workload 512k --steps 100 --warmup 10 --no-cpu-baseline --no-configs --e2e-steps 3 > gpurun_out/r02v_512k.json 2> gpurun_out/r02v_512k.err; python tools/benchrow.py 512k gpurun_out/r02v_512k.json
