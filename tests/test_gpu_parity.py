"""GPU parity tests: every call goes through the C ABI (libpatchy_b200.so) and is compared with
the CPU oracle (oracle/pp_oracle.c, pinned against the unmodified reference) and with the golden
vectors generated from the reference itself (tools/make_golden.py).

Bar: bit-exact integer bond counts, overlap flags, deltaE and accept/reject decisions (FP64
threshold tests, MC.c:232-266, 328-335); bit-exact positions; orientations within 4e-15 after
sweeps (CUDA's sincos and glibc's sin/cos may differ in the last bit).
"""
import os

import numpy as np
import pytest

from conftest import ROOT, load_golden, replay_box_moves

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pp():
    import patchyparticles_b200 as pkg
    pkg.load_library()
    return pkg


def make_engine(pp, params, **kw):
    return pp.Engine(kf_delta=float(params[0]), kf_cosmax=float(params[1]), temperature=float(params[2]), **kw)


@pytest.mark.parametrize("name", ["fcc", "bcc", "gen1000_rho0.1", "gen2000_rho0.5"])
def test_total_bonds_golden(pp, name):
    g = load_golden(name)
    for tag in ("a", "b"):
        if f"energy_{tag}" not in g:
            continue
        e = make_engine(pp, g[f"params_{tag}"])
        e.upload(g["box"], g["r"], g["orient"])
        bonds, ov = e.total_bonds()
        assert -bonds == g[f"energy_{tag}"][0]
        assert ov == g[f"energy_{tag}"][1]
        assert np.array_equal(e.bond_counts(), g[f"bond_counts_{tag}"])
        e.close()


@pytest.mark.parametrize("name", ["fcc", "bcc", "gen1000_rho0.1", "gen2000_rho0.5"])
def test_proposals_frozen_golden(pp, name):
    g = load_golden(name)
    e = make_engine(pp, g["params_a"])
    e.upload(g["box"], g["r"], g["orient"])
    dE, ov, acc = e.eval_proposals(g["prop_idx"], g["prop_disp"], g["prop_orient"], g["prop_u"])
    assert np.array_equal(ov, g["prop_overlap"])
    assert np.array_equal(dE, g["prop_dE"])
    assert np.array_equal(acc, g["prop_accept"])
    e.close()


@pytest.mark.parametrize("name", ["fcc", "gen2000_rho0.5"])
def test_proposals_sequential_golden(pp, name):
    g = load_golden(name)
    n = len(g["seq_dE"])
    e = make_engine(pp, g["params_a"])
    e.upload(g["box"], g["r"], g["orient"])
    e.total_bonds()
    dE, ov, acc = e.eval_proposals(g["prop_idx"][:n], g["prop_disp"][:n], g["prop_orient"][:n], g["prop_u"][:n], commit=True)
    assert np.array_equal(ov, g["seq_overlap"])
    assert np.array_equal(dE, g["seq_dE"])
    assert np.array_equal(acc, g["seq_accept"])
    r, m = e.download()
    assert np.array_equal(r, g["seq_r"])
    assert np.array_equal(m, g["seq_orient"])
    assert e.observables().energy == g["seq_energy"][0]
    bonds, _ = e.total_bonds()
    assert -bonds == g["seq_energy"][0]
    e.close()


def test_roundtrip_and_edges(pp, oracle):
    # upload -> download is the identity in the caller's particle order
    box, r, m = pp.random_fluid(3000, 0.4, seed=3)
    e = pp.Engine()
    e.upload(box, r, m)
    r2, m2 = e.download()
    assert np.array_equal(r, r2) and np.array_equal(m, m2)
    # empty system
    e.upload(box, np.zeros((0, 3)), np.zeros((0, 3, 3)))
    assert e.total_bonds() == (0, 0)
    e.sweep(2)
    # single particle
    e.upload(box, r[:1], m[:1])
    assert e.total_bonds() == (0, 0)
    e.sweep(3)
    assert e.observables().tries[0] == 3
    # overlapping pair is flagged
    rr = r[:2].copy()
    rr[1] = rr[0] + np.array([0.5, 0.1, 0.0])
    e.upload(box, rr, m[:2])
    assert e.total_bonds()[1] == 1
    # smallest supported box: two cells per side, periodic images handled once
    small = np.array([2.5, 2.5, 2.5])
    rs = np.array([[0.1, 0.1, 0.1], [1.2, 0.1, 0.1], [0.1, 1.3, 2.4], [2.45, 1.25, 1.2]])
    ms = m[:4]
    e.upload(small, rs, ms)
    port = oracle.PortSystem(small, rs, ms)
    assert e.total_bonds() == port.total_bonds()
    # box too small for the checkerboard is refused with a message
    with pytest.raises(Exception):
        e.upload(np.array([2.0, 5.0, 5.0]), rs, ms)
    e.close()


@pytest.mark.parametrize("N,rho,delta,cosmax", [(2000, 0.5, 0.119, 0.98), (1500, 0.3, 0.25, 0.8), (4000, 0.6, 0.119, 0.92)])
def test_total_bonds_vs_port(pp, oracle, N, rho, delta, cosmax):
    box, r, m = pp.random_fluid(N, rho, seed=N)
    box = box * np.array([1.0, 1.1, 1.05])  # anisotropic box (stretching cannot create overlaps)
    r = r * np.array([1.0, 1.1, 1.05])
    port = oracle.PortSystem(box, r, m, kf_delta=delta, kf_cosmax=cosmax)
    e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax)
    e.upload(box, r, m)
    assert e.total_bonds() == port.total_bonds()
    assert np.array_equal(e.bond_counts(), port.bond_counts()[0])
    e.close()


@pytest.mark.parametrize("name", ["fcc", "gen2000_rho0.5"])
def test_bond_lists_golden(pp, oracle, name):
    """pp_bond_lists (what output_print_bonds lists, output.c:213-268): the sets of bonded neighbours equal the golden
    vectors (the reference's MC_interact over all pairs) and the restatement's; truncation keeps the counts."""
    g = load_golden(name)
    b = np.load(os.path.join(ROOT, "tests", "golden", "bond_lists.npz"))
    delta, cosmax = b[f"{name}_params"]
    e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax)
    e.upload(g["box"], g["r"], g["orient"])
    counts, nb = e.bond_lists(12)
    assert np.array_equal(counts, b[f"{name}_counts"])
    assert np.array_equal(oracle.sorted_bond_lists(nb), b[f"{name}_neighbours"])
    port = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax)
    pc, pnb = port.bond_lists(12)
    assert np.array_equal(counts, pc) and np.array_equal(oracle.sorted_bond_lists(nb), oracle.sorted_bond_lists(pnb))
    c2, nb2 = e.bond_lists(2)  # more bonds than columns: the counts stay, the lists hold the first two
    assert np.array_equal(c2, counts) and np.array_equal(nb2, nb[:, :2])
    assert np.array_equal(e.bond_counts(), counts)
    e.sweep(3)  # after moves (cell-sorted storage reshuffled): still consistent with a fresh count
    counts, nb = e.bond_lists(12)
    r, m = e.download()
    port.set_state(r, m)
    pc, pnb = port.bond_lists(12)
    assert np.array_equal(counts, pc) and np.array_equal(oracle.sorted_bond_lists(nb), oracle.sorted_bond_lists(pnb))
    e.close()


@pytest.mark.parametrize("lanes,cell_lanes,cmo,mid", [(0, 0, -1, -1), (0, 1, 1, -1), (0, 2, 2, -1), (0, 1, 0, -1), (1, 1, -1, -1), (1, 2, -1, -1),
                                                      (1, 4, -1, -1), (1, 8, -1, -1), (4, 0, -1, -1), (8, 0, -1, -1), (16, 0, -1, -1), (32, 0, -1, -1),
                                                      (8, 1, 3, -1), (0, 0, 1, 2), (0, 0, 1, 3), (0, 0, 2, 15), (0, 4, 1, 2)])
def test_sweep_matches_port(pp, oracle, lanes, cell_lanes, cmo, mid):
    """The CUDA sweep and the serial restatement of the same algorithm give the same trajectory, whichever
    kernel runs it: tile kernel of any width (lanes 4..32), thread-per-cell kernel with 1..8 lanes per cell, and the
    fused launch (cmo / mid: cells with at most cmo particles on one thread per cell, up to mid on 4 lanes per cell, the
    others on the 8-lane latency configuration; an explicit cell_lanes runs the three ranges as three launches)."""
    g = load_golden("gen2000_rho0.5")
    delta, cosmax, T = 0.25, 0.8, 0.15  # plenty of bonds
    port = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax, temperature=T, disp_max=0.2, theta_max=0.3, seed=99, lanes_per_cell=lanes, cell_lanes=cell_lanes,
                  cells_max_occupancy=cmo, cells_mid_occupancy=mid)
    e.upload(g["box"], g["r"], g["orient"])
    bonds, _ = e.total_bonds()
    pb, _ = port.total_bonds()
    assert bonds == pb
    st = oracle.SweepStats()
    for s in range(6):
        pb, st = port.sweep(99, s, 0.2, 0.3, pb, st)
        e.sweep(1)
        o = e.observables()
        assert -o.energy == pb, f"sweep {s}"
        assert o.tries[0] == st.tries and o.accepted[0] == st.accepted
        assert o.confined == st.confined and o.overlapped == st.overlapped
        assert o.pair_checks == st.pair_checks
        assert o.acceptance_roto == st.w_accepted / st.w_tries
    r, m = e.download()
    pr, pm = port.state()
    assert np.array_equal(r, pr)
    assert np.abs(m - pm).max() < 4e-15
    b2, ov, ok = e.check_energy()
    assert ok and b2 == pb and ov == 0
    e.close()


def test_sweep_invariants_large(pp):
    """MC_check_energy / cells_check equivalents on the bench-size system (N = 65536, rho = 0.5)."""
    box, r, m = pp.random_fluid(65536, 0.5, seed=11)
    e = pp.Engine(temperature=0.12, seed=5)
    e.upload(box, r, m)
    b0, ov = e.total_bonds()
    assert ov == 0
    e.sweep(20)
    o = e.observables()
    assert o.tries[0] == 20 * 65536
    b, ov, ok = e.check_energy()
    assert ok and ov == 0 and -o.energy == b
    r2, m2 = e.download()
    # orientations stay orthonormal, particles moved but not far
    assert np.abs(np.einsum("nij,nkj->nik", m2, m2) - np.eye(3)).max() < 1e-12
    assert 0 < np.abs(r2 - r).max() < 20 * 0.05 * 1.01
    e.close()


@pytest.mark.parametrize("lanes,cell_lanes,cmo,mid", [(0, 0, -1, -1), (0, 1, 3, -1), (0, 2, 1, -1), (1, 1, -1, -1), (1, 8, -1, -1), (8, 0, -1, -1),
                                                      (0, 0, 1, 2), (0, 0, 2, 3)])
def test_crystal_sweep_matches_port(pp, oracle, lanes, cell_lanes, cmo, mid):
    """BASELINE configs[4] in small: an FCC crystal at rho = 1.041 (dofccrrr.py recipe), 41-44 particles in a 27-cell
    neighbourhood, up to 4 per cell.  Dense staging areas, crowded-cell routing and the serial fallback all follow the
    CPU restatement bit for bit."""
    box, r, m = pp.fcc_configuration(1.041, 6, 6, 6)
    T = 1.0  # warm enough for ~1000 accepted moves in 8 sweeps, with displacements large enough for ~1000 overlaps
    port = oracle.PortSystem(box, r, m, temperature=T)
    e = pp.Engine(temperature=T, disp_max=0.15, theta_max=0.2, seed=5, lanes_per_cell=lanes, cell_lanes=cell_lanes, cells_max_occupancy=cmo,
                  cells_mid_occupancy=mid)
    e.upload(box, r, m)
    pb, pov = port.total_bonds()
    assert e.total_bonds() == (pb, pov) and pov == 0 and pb > 1000
    st = oracle.SweepStats()
    for s_ in range(8):
        pb, st = port.sweep(5, s_, 0.15, 0.2, pb, st)
    e.sweep(8)
    assert st.accepted > 200 and st.overlapped > 200
    o = e.observables()
    assert -o.energy == pb and o.tries[0] == st.tries and o.accepted[0] == st.accepted
    assert o.confined == st.confined and o.overlapped == st.overlapped and o.pair_checks == st.pair_checks
    rr, mm = e.download()
    pr, pm = port.state()
    assert np.array_equal(rr, pr) and np.abs(mm - pm).max() < 4e-15
    e.close()


def test_crystal_melt_invariants(pp):
    """BASELINE configs[4] at full size: FCC 40 x 40 x 40 x 4 = 256 000 particles, rho = 1.041, T = 0.11, with the
    automatic kernel choice and with the hybrid routing forced."""
    box, r, m = pp.fcc_configuration(1.041, 40, 40, 40)
    results = []
    for cmo in (-1, 3):
        e = pp.Engine(temperature=0.11, seed=9, cells_max_occupancy=cmo)
        e.upload(box, r, m)
        b0, ov = e.total_bonds()
        assert ov == 0 and b0 > 400000  # nearly four bonds per particle in the perfect crystal
        e.sweep(12)
        o = e.observables()
        b, ov, ok = e.check_energy()
        assert ok and ov == 0 and -o.energy == b and o.tries[0] == 12 * 256000
        results.append((b, o.accepted[0], o.confined, o.overlapped, e.download()[0]))
        e.close()
    # the kernel choice does not change the trajectory
    assert results[0][:4] == results[1][:4] and np.array_equal(results[0][4], results[1][4])


def test_npt_and_lx_full_size_invariants(pp):
    """BASELINE configs[2] size (N = 32768, rho = 0.3): NPT sweeps with volume moves (Ensemble = 5, P = 0.5, T = 0.15)
    and NVT sweeps with Lx moves keep the invariants -- volume-move attempts are Binomial(N, 1/N) per step as in
    MC.c:60-68, the box volume moves (NPT) or is conserved with Ly = Lz (Lx), running energy == recount, no overlap."""
    N = 32768
    box, r, m = pp.random_fluid(N, 0.3, seed=5)
    e = pp.Engine(ensemble=5, temperature=0.15, P=0.5, rescale_factor_max=1e-3, seed=3)
    e.upload(box, r, m)
    e.total_bonds()
    e.sweep(60)
    o = e.observables()
    assert o.tries[0] == 60 * N                        # a sweep is N roto-translation attempts ...
    assert 20 <= o.tries[6] <= 120 and o.accepted[6] <= o.tries[6]  # ... plus Binomial(N, 1/N) volume moves per step
    V0, V1 = float(np.prod(box)), float(np.prod(o.box))
    assert (V1 != V0) == (o.accepted[6] > 0) and abs(V1 / V0 - 1) < 0.05
    b, ov, ok = e.check_energy()
    assert ok and ov == 0 and -o.energy == b
    r2, _ = e.download()
    assert np.isfinite(r2).all()
    e.close()
    L = float(box[0])
    e = pp.Engine(temperature=0.15, Lx_move=1, Lx_change_max=0.01, Lyz_min=20., Lyz_max=80., seed=4)
    e.upload(box, r, m)
    e.total_bonds()
    e.sweep(60)
    o = e.observables()
    assert o.tries[0] == 60 * N and 20 <= o.tries[7] <= 120
    assert abs(float(np.prod(o.box)) / L ** 3 - 1) < 1e-12 and o.box[1] == o.box[2]  # MC.c:521-524: V fixed, Ly = Lz
    b, ov, ok = e.check_energy()
    assert ok and ov == 0 and -o.energy == b
    e.close()


def test_full_size_invariants(pp):
    """BASELINE configs[3] size (N = 4 194 304, rho = 0.6) on one GPU, hybrid kernel routing: size-independent
    properties -- attempt count, running energy == recount (MC_check_energy), no overlap, bond counts sum to twice
    the bonds, orthonormal orientations, hard-core distances intact on a sample, download -> upload idempotence."""
    N = 4194304
    box, r, m = pp.random_fluid(N, 0.6, seed=3)
    e = pp.Engine(temperature=0.1, seed=17)
    e.upload(box, r, m)
    b0, ov = e.total_bonds()
    assert ov == 0
    e.sweep(9)  # three grid shifts
    o = e.observables()
    assert o.tries[0] == 9 * N and 0 < o.accepted[0] < o.tries[0]
    assert o.accepted[0] + o.confined + o.overlapped <= o.tries[0]
    b, ov, ok = e.check_energy()
    assert ok and ov == 0 and -o.energy == b
    assert int(e.bond_counts().astype(np.int64).sum()) == 2 * b
    r2, m2 = e.download()
    assert np.abs(np.einsum("nij,nkj->nik", m2[::64], m2[::64]) - np.eye(3)).max() < 1e-12
    assert 0 < np.abs(r2 - r).max() < 9 * 0.05 * 1.01
    e2 = pp.Engine(temperature=0.1, seed=17)
    e2.upload(box, r2, m2)
    assert e2.total_bonds() == (b, 0)  # idempotent: the downloaded state re-counts to the same bonds, no overlap
    e2.close()
    e.close()


def test_volume_move_vs_port(pp, oracle):
    g = load_golden("gen2000_rho0.5")
    delta, cosmax, T, P, rfm = 0.25, 0.8, 0.15, 0.5, 0.0005
    port = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    e = pp.Engine(ensemble=5, kf_delta=delta, kf_cosmax=cosmax, temperature=T, P=P, rescale_factor_max=rfm)
    e.upload(g["box"], g["r"], g["orient"])
    bonds, _ = e.total_bonds()
    pb, _ = port.total_bonds()
    rng = np.random.default_rng(5)
    n_acc = 0
    for k in range(40):
        u = rng.random(3)
        acc_p, pb, ov = port.volume_move(P, rfm, u[0], u[1], u[2], pb)
        acc_g = e.volume_move(u[0], u[1], u[2])
        assert acc_g == acc_p, f"move {k}"
        o = e.observables()
        assert -o.energy == pb
        assert np.array_equal(np.array(o.box), port.box())
        n_acc += acc_g
    assert 0 < n_acc < 40
    r, m = e.download()
    pr, pm = port.state()
    assert np.abs(r - pr).max() < 1e-12  # rejected moves: the reference divides back, the engine keeps the old bits
    assert e.observables().tries[6] == 40 and e.observables().accepted[6] == n_acc
    e.close()


def test_Lx_move_vs_port(pp, oracle):
    g = load_golden("gen2000_rho0.5")
    delta, cosmax, T = 0.25, 0.8, 0.15
    port = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    L = float(g["box"][1])
    e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax, temperature=T, Lx_move=1, Lx_change_max=0.01, Lyz_min=L - 0.002, Lyz_max=L + 0.002)
    e.upload(g["box"], g["r"], g["orient"])
    e.total_bonds()
    pb, _ = port.total_bonds()
    rng = np.random.default_rng(6)
    seen = set()
    for k in range(40):
        u = rng.random(2)
        acc_p, pb, ov = port.Lx_move(0.01, L - 0.002, L + 0.002, u[0], u[1], pb)
        acc_g = e.Lx_move(u[0], u[1])
        assert acc_g == acc_p, f"move {k}"
        assert -e.observables().energy == pb
        seen.add(acc_g)
    assert seen == {-1, 0, 1}
    e.close()


# ---- pins against the unmodified reference added in round 2 ------------------------------------------------------
def test_warm_bonded_golden(pp):
    """Bonded state equilibrated by the reference (delta 0.25, cosmax 0.8): its decisions on 8192 frozen and 4096 sequential
    proposals, hundreds accepted uphill through u < exp(-dE / T) (MC.c:332), bit for bit including the final state."""
    g = load_golden("warm1000")
    e = make_engine(pp, g["params"])
    e.upload(g["box"], g["r"], g["orient"])
    assert e.total_bonds() == (-g["energy"][0], 0)
    assert np.array_equal(e.bond_counts(), g["bond_counts"])
    dE, ov, acc = e.eval_proposals(g["prop_idx"], g["prop_disp"], g["prop_orient"], g["prop_u"])
    assert np.array_equal(ov, g["prop_overlap"]) and np.array_equal(dE, g["prop_dE"]) and np.array_equal(acc, g["prop_accept"])
    assert ((dE > 0) & (acc == 1)).sum() >= 300
    n = len(g["seq_dE"])
    dE, ov, acc = e.eval_proposals(g["prop_idx"][:n], g["prop_disp"][:n], g["prop_orient"][:n], g["prop_u"][:n], commit=True)
    assert np.array_equal(ov, g["seq_overlap"]) and np.array_equal(dE, g["seq_dE"]) and np.array_equal(acc, g["seq_accept"])
    r, m = e.download()
    assert np.array_equal(r, g["seq_r"]) and np.array_equal(m, g["seq_orient"])
    assert e.observables().energy == g["seq_energy"][0] and e.total_bonds() == (-g["seq_energy"][0], 0)
    e.close()


@pytest.mark.parametrize("tag", ["vol", "vol_grid", "lx", "lx_grid"])
def test_box_moves_golden(pp, tag):
    """pp_volume_move / pp_Lx_move against the reference's own MC_change_volume (MC.c:450-513) and MC_change_Lx
    (MC.c:515-585) driven with recorded uniforms: result, box (bit for bit) and energy after each of 60 moves, including
    moves that cross the re-grid threshold (MC.c:481-485) and Lx moves rejected by the Lyz window (MC.c:527-530)."""
    g = load_golden("box_moves")
    P, rfm, lx_max, lyz_min, lyz_max = g[f"{tag}_keys"]
    if tag.startswith("vol"):
        e = make_engine(pp, g[f"{tag}_params"], ensemble=5, P=float(P), rescale_factor_max=float(rfm))
    else:
        e = make_engine(pp, g[f"{tag}_params"], Lx_move=1, Lx_change_max=float(lx_max), Lyz_min=float(lyz_min), Lyz_max=float(lyz_max))
    e.upload(g[f"{tag}_box"], g[f"{tag}_r"], g[f"{tag}_orient"])
    assert e.total_bonds() == (-g[f"{tag}_energy0"][0], 0)

    def observe():
        o = e.observables()
        return np.array(o.box), o.energy, None

    replay_box_moves(g, tag, lambda P_, rfm_, a, b, c: e.volume_move(a, b, c), lambda mx, lo, hi, a, b: e.Lx_move(a, b), observe)
    n_acc = int((g[f"{tag}_result"] == 1).sum())
    o = e.observables()
    slot = 6 if tag.startswith("vol") else 7
    assert o.tries[slot] == len(g[f"{tag}_u"]) and o.accepted[slot] == n_acc
    r, _ = e.download()
    assert np.abs(r - g[f"{tag}_r_final"]).max() < 1e-12  # (a rejected move keeps the old bits here; the reference divides back)
    b, ov, ok = e.check_energy()
    assert ok and -b == g[f"{tag}_energies"][-1]
    e.sweep(3)  # and the engine goes on from the final box
    assert e.check_energy()[2]
    e.close()


def test_orthonormalise_vs_reference(pp):
    """K6 (k_orthonormalise inside pp_check_energy, MC.c:774) against the reference's gram_schmidt (utils.c:91-113) on
    drifted matrices: bit-equal."""
    g = load_golden("helpers")
    m_in, want = g["gs_in"], g["gs_out"]
    n = len(m_in)
    side = int(np.ceil(n ** (1 / 3)))
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
    box = np.full(3, 3.0 * side)
    e = pp.Engine()
    e.upload(box, 3.0 * grid + 0.7, m_in)  # isolated particles: the energy does not care about their orientation
    assert e.total_bonds() == (0, 0)
    e.check_energy()
    _, m = e.download()
    assert np.array_equal(m, want)
    e.close()


def test_box_move_into_too_small_box_is_rejected(pp):
    """A trial box that cannot hold two cells per side is counted and rejected with nothing touched (the reference clamps
    its grid to 3 cells there, cells.c:24-27); the run goes on."""
    rng = np.random.default_rng(2)
    box = np.array([2.3, 6.0, 6.0])  # two cells of 1.15 > r_cut = 1.119 along x
    r = np.array([[0.2, 0.3, 0.4], [1.5, 2.0, 2.5], [0.3, 4.0, 1.0], [1.4, 5.0, 4.5]])
    m = pp.random_fluid(64, 0.2, seed=1)[2][:4]
    e = pp.Engine(ensemble=5, P=50.0, rescale_factor_max=0.5, temperature=1.0)
    e.upload(box, r, m)
    e.total_bonds()
    acc = e.volume_move(0.1, 0.0, 1e-9)  # compress x by exp(-0.25): 1.79 < 2 r_cut
    o = e.observables()
    assert acc == 0 and np.array_equal(np.array(o.box), box) and o.tries[6] == 1 and o.accepted[6] == 0
    r2, _ = e.download()
    assert np.array_equal(r2, r)
    e.sweep(20)  # NPT steps with more such trials in them
    assert e.check_energy()[2]
    e.close()


def test_gc_moves_golden(pp):
    """MC_add_remove (MC.c:342-448): the engine, fed the recorded uniforms of tests/golden/gc_moves.npz, takes the decisions the
    reference's own MC_add_remove took (400 moves, insertions and removals, 49 accepted), consumes the same number of uniforms,
    follows the same N and running energy, and ends in the same configuration -- index for index (a new particle is N, the
    last particle takes the index of a removed one), positions and orientation matrices bit for bit."""
    g = load_golden("gc_moves")
    delta, cosmax, T, z, n_min, n_max = g["params"]
    e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax, temperature=T, seed=1)
    e.reserve(int(n_max))
    e.upload(g["box"], g["r"], g["orient"])
    bonds, ov = e.total_bonds()
    assert ov == 0 and -bonds == g["energy0"][0]
    # single-particle energies against the reference's per-particle counts (MC_energy, MC.c:272-309)
    counts = e.bond_counts()
    for i in (0, 17, 512, 999):
        assert e.particle_bonds(i) == (counts[i], 0)
    for k in range(len(g["u"])):
        res = e.add_remove(z, int(n_min), int(n_max), g["u"][k])
        assert (res.kind, res.accepted, res.n_uniforms, res.N) == (g["kind"][k], g["accepted"][k], g["used"][k], g["N"][k]), f"move {k}"
        assert e.observables().energy == g["energies"][k], f"move {k}"
    r, m = e.download()
    assert np.array_equal(r, g["r_final"]) and np.array_equal(m, g["orient_final"])
    b2, ov2, ok = e.check_energy()
    assert ok and ov2 == 0 and -b2 == g["energies"][-1]
    e.sweep(3)  # the sweep runs on the edited particle set
    b3, ov3, ok3 = e.check_energy()
    assert ok3 and ov3 == 0
    e.close()


def test_bond_count_queue_overflow(pp, monkeypatch):
    """K3: when more pairs pass the FP32 screens than the pair queue holds, the count is repeated with the exact tile kernel
    (Counters.pad signals the overflow).  PP_BONDS_QUEUE_BYTES shrinks the queue to 511 pairs on a crystal with 900 bonds."""
    g = load_golden("fcc")
    delta, cosmax, T = g["params_a"]
    e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    e.upload(g["box"], g["r"], g["orient"])
    want = e.total_bonds()
    assert want == (-int(g["energy_a"][0]), 0) and want[0] > 600
    e.set_profiling(True)
    e.reset_kernel_stats()
    e.total_bonds()
    normal = e.kernel_stats()["bonds"][0]
    monkeypatch.setenv("PP_BONDS_QUEUE_BYTES", "4096")
    e.reset_kernel_stats()
    assert e.total_bonds() == want
    assert e.kernel_stats()["bonds"][0] == normal + 2  # the streaming launches, then the tile kernel's two
    monkeypatch.setenv("PP_BONDS_QUEUE_BYTES", "0")   # no queue at all: the tile kernel straight away
    assert e.total_bonds() == want
    e.close()


def test_neighbour_queries_golden(pp):
    """The neighbour queries of the cluster moves against the reference's own MC_interact over all pairs
    (tests/golden/neighbour_queries.npz): bonded neighbour per patch as AVB's _set_neighbours collects it (avb.c:43-75) and the
    particles within range as VMMC's _populate_possible_links meets them (vmmc.c:178)."""
    g = load_golden("neighbour_queries")
    delta, T = g["params"]
    N = len(g["r"])
    every = np.arange(N, dtype=np.int32)
    for tag, cosmax in (("narrow", 0.98), ("wide", 0.8)):
        e = pp.Engine(kf_delta=delta, kf_cosmax=cosmax, temperature=T)
        e.upload(g["box"], g["r"], g["orient"])
        nb, num = e.patch_neighbours(every)
        cnt, q, onp = g[f"{tag}_count"], g[f"{tag}_q"], g[f"{tag}_onp"]
        assert np.array_equal(num, cnt)
        for i in range(N):
            for patch in range(4):
                on_patch = [int(q[i, k]) for k in range(cnt[i]) if onp[i, k] == patch]
                if not on_patch:
                    assert nb[i, patch] == -1
                elif len(on_patch) == 1:
                    assert nb[i, patch] == on_patch[0]
                else:  # two neighbours on one patch: which one stays depends on the traversal order (avb.c:64); one of them
                    assert nb[i, patch] in on_patch
        # a subset in another order gives the same rows
        some = np.array([999, 3, 512, 3], dtype=np.int32)
        nb2, num2 = e.patch_neighbours(some)
        assert np.array_equal(nb2, nb[some]) and np.array_equal(num2, num[some])
        e.close()
    e = pp.Engine(kf_delta=delta, kf_cosmax=0.98, temperature=T)
    e.upload(g["box"], g["r"], g["orient"])
    cnt, nb = e.range_neighbours(every, 24)
    assert np.array_equal(cnt, g["range_count"])
    for i in range(N):
        assert sorted(nb[i, :cnt[i]].tolist()) == sorted(g["range_nb"][i, :cnt[i]].tolist()) and np.all(nb[i, cnt[i]:] == -1)
    cnt4, nb4 = e.range_neighbours(every, 4)  # a short list keeps the first ones in the same order
    assert np.array_equal(cnt4, cnt) and np.array_equal(nb4, nb[:, :4])
    e.close()
