#!/usr/bin/env python
"""Exercises every kernel class beside the sweep (K1 cell build, K3 bond count, K4 rescale, K6
re-orthonormalisation) once or twice on a synthetic fluid, so that ncu can list them:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --cache-control none \
        --clock-control none --csv --log-file gpurun_out/aux.csv python tools/aux_kernels.py 4194304 0.6

usage: aux_kernels.py N rho [n_volume_moves]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import patchyparticles_b200 as pkg  # noqa: E402

N, rho = int(sys.argv[1]), float(sys.argv[2])
n_vol = int(sys.argv[3]) if len(sys.argv) > 3 else 3
box, r, m = pkg.random_fluid(N, rho, seed=1)
eng = pkg.Engine(ensemble=pkg.engine.NPT, temperature=0.15, P=0.5, rescale_factor_max=1e-5, seed=7)
eng.upload(box, r, m)                       # K1
print("bonds", eng.total_bonds())           # K3
for k in range(n_vol):                      # K4 + K3 each
    print("volume move", k, eng.volume_move(0.1 + 0.3 * k, 0.5 + 0.1 * (k - 1), 0.5))
eng.sweep(2)                                # K1 (grid shift), K2a, K2b
print("check", eng.check_energy())          # K3 + K6
eng.close()
