"""GPU parity at the sizes BASELINE.json names, against the unmodified reference itself.

oracle/_ref/ (the reference compiled by oracle/Makefile: its generator and libppref.so = reference objects behind
oracle/ref_harness.c) travels to the GPU box with the repository snapshot, so these tests run the reference's own
generator for the BASELINE configurations and compare the CUDA path with the reference's MC_energy / MC_interact /
MC.c:328-335 on the very same configuration: bond count, per-particle counts and >= 10^5 frozen proposals, bit-exact.
"""
import os

import numpy as np
import pytest

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pp():
    import patchyparticles_b200 as pkg
    pkg.load_library()
    return pkg


def rotations(rng, m_old, theta_max):
    """new = old . R(axis, theta) for uniformly random axes (vectorised; the evaluators take the matrix as an input)"""
    n = len(m_old)
    axis = rng.normal(size=(n, 3))
    axis /= np.linalg.norm(axis, axis=1)[:, None]
    th = rng.random(n) * theta_max
    K = np.zeros((n, 3, 3))
    K[:, 0, 1], K[:, 0, 2], K[:, 1, 0], K[:, 1, 2], K[:, 2, 0], K[:, 2, 1] = -axis[:, 2], axis[:, 1], axis[:, 2], -axis[:, 0], -axis[:, 1], axis[:, 0]
    R = np.eye(3)[None] + np.sin(th)[:, None, None] * K + (1 - np.cos(th))[:, None, None] * (K @ K)
    return m_old @ R


def proposal_list(rng, N, m, n, disp_max, theta_max):
    idx = rng.integers(0, N, n).astype(np.int32)
    return idx, (rng.random((n, 3)) - 0.5) * disp_max, rotations(rng, m[idx], theta_max), rng.random(n)


def write_exact_rrr(path, box, r, m):
    """17 significant digits: the reference parses back the very same doubles (rows 0 and 1; it rebuilds row 2)"""
    with open(path, "w") as f:
        f.write("0 %d %.17g %.17g %.17g\n" % (len(r), *box))
        for i in range(len(r)):
            f.write("%.17g %.17g %.17g\n%.17g %.17g %.17g\n%.17g %.17g %.17g\n" % (*m[i, 0], *m[i, 1], *r[i]))


@pytest.mark.parametrize("N,rho,bonds,n_prop", [(65536, 0.5, 59, 120000), (32768, 0.3, 15, 30000)])
def test_generator_configurations_vs_reference(pp, oracle, tmp_path, N, rho, bonds, n_prop):
    """BASELINE configs[1] / configs[2]: the reference generator's own output (deterministic, generator.c:86-124)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    rrr = str(tmp_path / "generated.rrr")
    oracle.run_generator(N, rho, rrr)
    known = load_golden("generator_known")[f"{N}_{rho}"]
    ref = oracle.RefSystem(rrr)
    try:
        box, (r, m) = ref.box(), ref.state()
        assert float(np.sum(r)) == known[2]  # the same configuration the golden file was made from
        E, ov = ref.total_energy()
        assert (E, ov) == (-bonds, 0) == (known[0], known[1])
        counts, _ = ref.bond_counts()
        e = pp.Engine()
        e.upload(box, r, m)
        assert e.total_bonds() == (bonds, 0)
        assert np.array_equal(e.bond_counts(), counts)
        rng = np.random.default_rng(N)
        idx, disp, on, u = proposal_list(rng, N, m, n_prop, 0.4, 0.5)
        # half of the proposals on bonded particles and their partners, so that dE != 0 is exercised at this size too
        bonded = np.flatnonzero(counts > 0).astype(np.int32)
        idx[::2] = bonded[rng.integers(0, len(bonded), len(idx[::2]))]
        on[::2] = rotations(rng, m[idx[::2]], 0.5)
        want = ref.eval_proposals(idx, disp, on, u)
        got = e.eval_proposals(idx, disp, on, u)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        assert (want[0] != 0).sum() > 100 and (want[1] == 1).sum() > 1000 and (want[2] == 1).sum() > 1000
        e.close()
    finally:
        ref.close()


def test_fcc_256k_vs_port(pp, oracle):
    """BASELINE configs[4]: FCC 40 x 40 x 40 x 4 = 256 000 (dofccrrr.py layout): bonds and per-particle counts against the
    restatement, which the CPU suite pins on the reference's fcc.rrr (-914)."""
    box, r, m = pp.fcc_configuration(1.041, 40, 40, 40)
    port = oracle.PortSystem(box, r, m, kf_cosmax=0.92)
    e = pp.Engine(kf_cosmax=0.92)
    e.upload(box, r, m)
    pb = port.total_bonds()
    assert e.total_bonds() == pb and pb[0] > 400000 and pb[1] == 0
    assert np.array_equal(e.bond_counts(), port.bond_counts()[0])
    e.close()


def test_wide_box_vs_reference(pp, oracle, tmp_path):
    """Coordinates of the size they have at N = 4M (box 191.2, cells 170^3), where rint(d * (1 / L)) against rint(d / L)
    and the division-free cell index matter most: a dense droplet straddling the periodic boundary of a 191.2066 box,
    positions unwrapped beyond the box like the reference keeps them.  K3, per-particle counts and K5 against the
    reference's own functions."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not present")
    small, r, m = pp.random_fluid(6000, 0.6, seed=5)
    L = 191.2066
    box = np.array([L, L, L])
    r = r + np.array([L - 0.5 * small[0], 170.0, L - 3.0])  # x and z cross the boundary, |r| up to 200
    delta, cosmax, T = 0.25, 0.8, 0.2
    rrr = str(tmp_path / "wide.rrr")
    write_exact_rrr(rrr, box, r, m)
    ref = oracle.RefSystem(rrr, KF_delta=delta, KF_cosmax=cosmax, Temperature=T)
    try:
        rr, mm = ref.state()
        assert np.array_equal(rr, r)
        E, ov = ref.total_energy()
        counts, _ = ref.bond_counts()
        e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax, temperature=T)
        e.upload(box, rr, mm)
        assert e.total_bonds() == (-E, ov) and E < -1000
        assert np.array_equal(e.bond_counts(), counts)
        rng = np.random.default_rng(9)
        idx, disp, on, u = proposal_list(rng, len(r), mm, 40000, 0.4, 0.5)
        want = ref.eval_proposals(idx, disp, on, u)
        got = e.eval_proposals(idx, disp, on, u)
        for a, b in zip(got, want):
            assert np.array_equal(a, b)
        # the sweep kernels on the same coordinates: bit-identical to the serial restatement, then a reference recount
        port = oracle.PortSystem(box, rr, mm, kf_delta=delta, kf_cosmax=cosmax, temperature=T)
        e2 = pp.Engine(kf_delta=delta, kf_cosmax=cosmax, temperature=T, disp_max=0.2, theta_max=0.3, seed=3, lanes_per_cell=1)
        e2.upload(box, rr, mm)
        pb = e2.total_bonds()[0]
        st = oracle.SweepStats()
        for s in range(3):
            pb, st = port.sweep(3, s, 0.2, 0.3, pb, st)
        e2.sweep(3)
        o = e2.observables()
        assert -o.energy == pb and o.accepted[0] == st.accepted
        r3, m3 = e2.download()
        assert np.array_equal(r3, port.state()[0])
        ref.set_state(r3, m3)
        assert ref.total_energy() == (-float(pb), 0)
        e.close()
        e2.close()
    finally:
        ref.close()
