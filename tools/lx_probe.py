#!/usr/bin/env python
"""Runs the reference program at the state point of tests/test_host_and_statistics.py::test_lx_statistics_vs_reference and
prints block averages and the autocorrelation of Lx (used to choose a state point whose error bars are honest).

usage: lx_probe.py [binary] [steps]"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
binary = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "oracle", "_ref", "PatchyParticles")
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
d = tempfile.mkdtemp(prefix="lx_probe_")
subprocess.run([os.path.join(ROOT, "oracle", "_ref", "generator"), "300", "0.25"], cwd=d, check=True, stderr=subprocess.DEVNULL)
keys = dict(Dynamics=0, Ensemble=0, Temperature=0.3, Steps=steps, Disp_max=0.2, Theta_max=0.3, Initial_conditions_file="generated.rrr",
            Print_every=50, Save_every=100000000, Configuration_folder=".", KF_delta=0.119, KF_cosmax=0.92, Seed=12345, Log_file="none",
            Lx_move=1, Lx_change_max=0.4, Lyz_min=10.45, Lyz_max=10.8)
with open(os.path.join(d, "input"), "w") as f:
    for k, v in keys.items():
        f.write(f"{k} = {v}\n")
subprocess.run([binary, "input"], cwd=d, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
b = np.loadtxt(os.path.join(d, "boxshape.dat"))
a = np.loadtxt(os.path.join(d, "acceptance.dat"))
e = np.loadtxt(os.path.join(d, "energy.dat"))
b = b[b[:, 0] >= 3000]


def bs(x, n=10):
    x = np.asarray(x)
    m = (len(x) // n) * n
    mm = x[:m].reshape(n, -1).mean(1)
    return mm.mean(), mm.std(ddof=1) / np.sqrt(n)


print("Lx", bs(b[:, 1]), "var", bs((b[:, 1] - b[:, 1].mean()) ** 2), "acc", bs(a[a[:, 0] >= 3000, 2]), "E/N", bs(e[e[:, 0] >= 3000, 1]))
print("range", b[:, 1].min(), b[:, 1].max(), "Ly", b[:, 2].min(), b[:, 2].max())
x = b[:, 1] - b[:, 1].mean()
print("autocorrelation at lags (x50 steps)", [(k, round(float((x[:-k] * x[k:]).mean() / x.var()), 3)) for k in (1, 2, 5, 10, 20)])
