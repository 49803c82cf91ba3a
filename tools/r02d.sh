#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest.log 2>&1; tail -n 40 gpurun_out/r02d_pytest.log
