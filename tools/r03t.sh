#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03t_smoke.log 2>&1; tail -n 3 gpurun_out/r03t_smoke.log
