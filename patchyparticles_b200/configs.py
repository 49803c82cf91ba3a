"""Synthetic configurations for tests and benchmarks.

``fcc_configuration`` restates the lattice recipe of the reference's Python-2 helper
dofccrrr.py:22-85 (4-site FCC cell, alternating +-45 degree rotations about z, all coordinates
shifted by 0.01); ``random_fluid`` is a seeded hard-sphere fluid built by thinning a jittered
lattice (a fast stand-in for generator.c:43-84 at sizes where random sequential insertion takes
minutes).
"""
import numpy as np


def fcc_configuration(rho, nx, ny, nz):
    a = (4.0 / rho) ** (1.0 / 3.0)
    sites = np.array([[0., 0., 0.], [.5, .5, 0.], [.5, 0., .5], [0., .5, .5]])
    l, m, n = np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij")
    cells = np.stack([l.ravel(), m.ravel(), n.ravel()], axis=1).astype(np.float64)
    r = ((cells[:, None, :] + sites[None, :, :]) * a).reshape(-1, 3) + 0.01
    N = r.shape[0]
    orient = np.zeros((N, 3, 3))
    for k, al in enumerate([np.pi / 4., np.pi / 4., -np.pi / 4., -np.pi / 4.]):
        ct, st = np.cos(al), np.sin(al)
        orient[k::4] = np.array([[ct, -st, 0.], [st, ct, 0.], [0., 0., 1.]])
    box = np.array([nx * a, ny * a, nz * a])
    return box, r, orient


def random_orientations(n, rng):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1)[:, None]
    w, x, y, z = q.T
    m = np.empty((n, 3, 3))
    m[:, 0, 0] = 1 - 2 * (y * y + z * z)
    m[:, 0, 1] = 2 * (x * y - z * w)
    m[:, 0, 2] = 2 * (x * z + y * w)
    m[:, 1, 0] = 2 * (x * y + z * w)
    m[:, 1, 1] = 1 - 2 * (x * x + z * z)
    m[:, 1, 2] = 2 * (y * z - x * w)
    m[:, 2, 0] = 2 * (x * z - y * w)
    m[:, 2, 1] = 2 * (y * z + x * w)
    m[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return m


def random_fluid(N, rho, seed=0):
    """N hard spheres (diameter 1) at number density rho in a cubic box, no overlaps, random
    orientations.  Simple-cubic lattice with spacing >= 1 + jitter margin, randomly thinned to N."""
    rng = np.random.default_rng(seed)
    L = (N / rho) ** (1.0 / 3.0)
    n_side = int(np.floor(L / 1.0000001))
    while n_side ** 3 < N:
        raise ValueError("density too high for the lattice-thinning generator")
    a = L / n_side
    jitter = 0.5 * (a - 1.0) * 0.999
    idx = rng.choice(n_side ** 3, size=N, replace=False)
    ix, iy, iz = np.unravel_index(idx, (n_side, n_side, n_side))
    r = (np.stack([ix, iy, iz], axis=1) + 0.5) * a + (rng.random((N, 3)) * 2 - 1) * jitter
    return np.array([L, L, L]), r, random_orientations(N, rng)
