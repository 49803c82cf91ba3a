#!/usr/bin/env python
"""A small run through every kernel class, for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python tools/sanitize_run.py

N = 2000 at rho = 0.5 with plenty of bonds (delta = 0.25, cosmax = 0.8, T = 0.15): the tile kernel at every width, the
thread-per-cell kernel with 1 .. 8 lanes, the fused launch, K1 (grid shifts), K3 (both kernels), K4 + box moves, K5, K6,
bond lists, and two virtual slabs on one device (halo mirror stores, phase kernels, migration mailboxes)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import patchyparticles_b200 as pkg  # noqa: E402

box, r, m = pkg.random_fluid(2000, 0.5, seed=3)
configs = [dict(), dict(lanes_per_cell=1, cell_lanes=1), dict(lanes_per_cell=1, cell_lanes=2), dict(lanes_per_cell=1, cell_lanes=4),
           dict(lanes_per_cell=1, cell_lanes=8), dict(lanes_per_cell=4), dict(lanes_per_cell=8), dict(lanes_per_cell=16), dict(lanes_per_cell=32),
           dict(cells_max_occupancy=1, cells_mid_occupancy=2), dict(cells_max_occupancy=2, cells_mid_occupancy=15),
           dict(cell_lanes=4, cells_max_occupancy=1, cells_mid_occupancy=2)]
for kw in configs:
    e = pkg.Engine(kf_delta=0.25, kf_cosmax=0.8, temperature=0.15, disp_max=0.2, theta_max=0.3, seed=99, **kw)
    e.upload(box, r, m)
    b0, _ = e.total_bonds()
    e.sweep(5)
    b1, ov, ok = e.check_energy()
    assert ok and ov == 0
    print(kw, "bonds", b0, "->", b1)
    e.close()
e = pkg.Engine(ensemble=pkg.engine.NPT, kf_delta=0.25, kf_cosmax=0.8, temperature=0.3, P=0.3, rescale_factor_max=0.02, seed=5)
e.upload(box, r, m)
e.total_bonds()
e.sweep(6)
print("npt", e.check_energy(), e.volume_move(0.3, 0.7, 0.5))
counts, nb = e.bond_lists(12)
rng = np.random.default_rng(0)
idx = rng.integers(0, 2000, 64).astype(np.int32)
print("proposals", sum(e.eval_proposals(idx, (rng.random((64, 3)) - 0.5) * 0.3, m[idx], rng.random(64))[2]), "bond lists", int(counts.sum()))
e.close()
g = pkg.Group(2, devices=[0, 0], kf_delta=0.25, kf_cosmax=0.8, temperature=0.15, disp_max=0.2, theta_max=0.3, seed=99)
g.upload(box, r, m)
g.sweep(9)
print("group", g.check_energy())
g.close()
print("sanitize_run ok")
