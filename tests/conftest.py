import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

GOLDEN = os.path.join(HERE, "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


@pytest.fixture(scope="session")
def oracle():
    import _oracle
    _oracle.build_oracle()
    return _oracle


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def replay_box_moves(g, tag, volume_move, Lx_move, observe):
    """Replays one golden sequence of tests/golden/box_moves.npz through the callables of an implementation and compares
    every move with what the reference's MC_change_volume / MC_change_Lx did (result, box bit for bit, energy)."""
    P, rfm, lx_max, lyz_min, lyz_max = g[f"{tag}_keys"]
    u, res, used = g[f"{tag}_u"], g[f"{tag}_result"], g[f"{tag}_used"]
    for k in range(len(u)):
        if tag.startswith("vol"):
            got = volume_move(P, rfm, u[k, 0], u[k, 1], u[k, 2])
        else:
            got = Lx_move(lx_max, lyz_min, lyz_max, u[k, 0], u[k, 1])
        box, energy, overlap = observe()
        assert got == res[k], f"{tag} move {k}: result {got}, the reference's {res[k]}"
        assert np.array_equal(box, g[f"{tag}_boxes"][k]), f"{tag} move {k}: box"
        assert energy == g[f"{tag}_energies"][k], f"{tag} move {k}: energy"
        if overlap is not None and res[k] >= 0:  # the reference skips its acceptance draw exactly when it found an overlap
            assert bool(overlap) == (used[k] == (2 if tag.startswith("vol") else 1)), f"{tag} move {k}: overlap"
