"""CPU tests (no GPU): the oracle restatement against the golden vectors generated from the
unmodified reference, against the reference library itself when it has been built
(oracle/_ref/libppref.so), the host-side helpers, and the C-ABI surface of the product library."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_golden, replay_box_moves

GOLDEN_SETS = ["fcc", "bcc", "gen1000_rho0.1", "gen2000_rho0.5"]


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_port_total_bonds_golden(oracle, name):
    g = load_golden(name)
    for tag in ("a", "b"):
        if f"energy_{tag}" not in g:
            continue
        delta, cosmax, T = g[f"params_{tag}"]
        p = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
        bonds, ov = p.total_bonds()
        assert -bonds == g[f"energy_{tag}"][0] and ov == g[f"energy_{tag}"][1]
        assert np.array_equal(p.bond_counts()[0], g[f"bond_counts_{tag}"])
        assert np.array_equal(p.cells(), g[f"ref_cells_{tag}"])


@pytest.mark.parametrize("name", ["fcc", "gen2000_rho0.5"])
def test_port_bond_lists_golden(oracle, name):
    """Bonded-neighbour lists (output.c:213-268) of the restatement against the golden vectors, which are the
    reference's own MC_interact over all pairs; live against the reference when it is built."""
    g = load_golden(name)
    b = np.load(os.path.join(ROOT, "tests", "golden", "bond_lists.npz"))
    delta, cosmax = b[f"{name}_params"]
    p = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax)
    counts, nb = p.bond_lists(12)
    assert np.array_equal(counts, b[f"{name}_counts"])
    assert np.array_equal(oracle.sorted_bond_lists(nb), b[f"{name}_neighbours"])
    assert np.array_equal((nb >= 0).sum(axis=1), counts)
    # a bond is mutual
    pairs = {(i, int(q)) for i in range(len(counts)) for q in nb[i] if q >= 0}
    assert all((q, i) in pairs for i, q in pairs)
    assert len(pairs) == 2 * p.total_bonds()[0]


GEN_BIN = os.path.join(ROOT, "host", "patchy_generate")


def test_generator_at_scale(oracle, tmp_path):
    """host/patchy_generate: the files load like the reference generator's (generator.c:73-81
    layout), hold N non-overlapping spheres at the requested density, are reproducible per seed, and the binary
    side format round-trips bit for bit.  The reference itself accepts them (no overlap, same energy)."""
    import subprocess
    import patchyparticles_b200 as pkg
    if not os.path.exists(GEN_BIN):
        pytest.skip("host/patchy_generate not built")
    run = lambda *a: subprocess.run([GEN_BIN, *[str(x) for x in a]], cwd=tmp_path, check=True, capture_output=True)
    run(3000, 0.6, 7, "a.rrr")
    run(3000, 0.6, 7, "b.rrr")
    run(3000, 0.6, 8, "c.rrr")
    run(3000, 0.6, 7, "a.rrb")
    assert open(tmp_path / "a.rrr").read() == open(tmp_path / "b.rrr").read() != open(tmp_path / "c.rrr").read()
    step, box, r, m = pkg.read_rrr(str(tmp_path / "a.rrr"))
    assert step == 0 and len(r) == 3000 and np.allclose(box, (3000 / 0.6) ** (1 / 3), rtol=0, atol=1e-11)
    assert r.min() >= 0 and r.max() <= box[0]
    assert np.abs(np.einsum("nij,nkj->nik", m, m) - np.eye(3)).max() < 1e-12
    port = oracle.PortSystem(box, r, m, kf_delta=0.25, kf_cosmax=0.8)
    bonds, overlap = port.total_bonds()
    assert overlap == 0 and bonds > 0  # random orientations at rho = 0.6: a few bonds
    # binary twin: same state at full precision (the text file rounds positions to 1e-12, rows to 1e-6)
    sb, bb, rb, mb = pkg.read_rrr(str(tmp_path / "a.rrb"))
    assert sb == 0 and np.abs(bb - box).max() < 1e-11 and np.abs(rb - r).max() < 1e-11 and np.abs(mb - m).max() < 2e-6
    pkg.write_rrb(str(tmp_path / "again.rrb"), 5, bb, rb, mb)
    s2, b2, r2, m2 = pkg.read_rrb(str(tmp_path / "again.rrb"))
    assert s2 == 5 and np.array_equal(b2, bb) and np.array_equal(r2, rb) and np.array_equal(m2, mb)
    # densities the reference generator refuses (generator.c:99-102) still work, impossible ones are refused
    run(4000, 0.9, 1, "dense.rrr")
    _, box_d, r_d, m_d = pkg.read_rrr(str(tmp_path / "dense.rrr"))
    assert oracle.PortSystem(box_d, r_d, m_d).total_bonds()[1] == 0
    p = subprocess.run([GEN_BIN, "4000", "1.2"], cwd=tmp_path, capture_output=True, text=True)
    assert p.returncode == 1 and "cannot produce" in p.stderr
    if oracle.have_ref():
        ref = oracle.RefSystem(str(tmp_path / "a.rrr"), Temperature=0.1, KF_delta=0.25, KF_cosmax=0.8)
        e, ov = ref.total_energy()
        assert ov == 0 and -e == bonds
        ref.close()


def test_known_answers_from_survey(oracle):
    """fcc.rrr -> -914 (343x4, 143x3, 13x2, 1x1); bcc.rrr -> 0 at delta 0.119, -92 at 0.135."""
    g = load_golden("fcc")
    assert g["energy_a"][0] == -914 and g["energy_b"][0] == -914
    assert list(np.bincount(g["bond_counts_a"])) == [0, 1, 13, 143, 343]
    g = load_golden("bcc")
    assert g["energy_a"][0] == 0 and g["energy_b"][0] == -92
    k = np.load(os.path.join(ROOT, "tests", "golden", "generator_known.npz"))
    assert k["32768_0.3"][0] == -15 and k["65536_0.5"][0] == -59


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_port_proposals_golden(oracle, name):
    g = load_golden(name)
    delta, cosmax, T = g["params_a"]
    p = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    dE, ov, acc = p.eval_proposals(g["prop_idx"], g["prop_disp"], g["prop_orient"], g["prop_u"])
    assert np.array_equal(dE, g["prop_dE"]) and np.array_equal(ov, g["prop_overlap"]) and np.array_equal(acc, g["prop_accept"])
    n = len(g["seq_dE"])
    dE, ov, acc = p.eval_proposals(g["prop_idx"][:n], g["prop_disp"][:n], g["prop_orient"][:n], g["prop_u"][:n], commit=True)
    assert np.array_equal(dE, g["seq_dE"]) and np.array_equal(ov, g["seq_overlap"]) and np.array_equal(acc, g["seq_accept"])
    r, m = p.state()
    assert np.array_equal(r, g["seq_r"]) and np.array_equal(m, g["seq_orient"])
    assert -p.total_bonds()[0] == g["seq_energy"][0]


def test_port_warm_bonded_golden(oracle):
    """A bonded state point equilibrated by the reference (delta 0.25, cosmax 0.8): the reference's decisions on 8192 frozen
    and 4096 sequential proposals, hundreds of them accepted uphill through u < exp(-dE / T) (MC.c:332)."""
    g = load_golden("warm1000")
    delta, cosmax, T = g["params"]
    uphill = (g["prop_dE"] > 0) & (g["prop_accept"] == 1)
    assert uphill.sum() >= 300 and ((g["seq_dE"] > 0) & (g["seq_accept"] == 1)).sum() >= 150  # the golden set does exercise the draw
    assert ((g["prop_dE"] > 0) & (g["prop_accept"] == 0) & (g["prop_overlap"] == 0)).sum() >= 1000
    p = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    assert p.total_bonds() == (-g["energy"][0], 0)
    assert np.array_equal(p.bond_counts()[0], g["bond_counts"])
    dE, ov, acc = p.eval_proposals(g["prop_idx"], g["prop_disp"], g["prop_orient"], g["prop_u"])
    assert np.array_equal(dE, g["prop_dE"]) and np.array_equal(ov, g["prop_overlap"]) and np.array_equal(acc, g["prop_accept"])
    n = len(g["seq_dE"])
    dE, ov, acc = p.eval_proposals(g["prop_idx"][:n], g["prop_disp"][:n], g["prop_orient"][:n], g["prop_u"][:n], commit=True)
    assert np.array_equal(dE, g["seq_dE"]) and np.array_equal(ov, g["seq_overlap"]) and np.array_equal(acc, g["seq_accept"])
    r, m = p.state()
    assert np.array_equal(r, g["seq_r"]) and np.array_equal(m, g["seq_orient"])
    assert -p.total_bonds()[0] == g["seq_energy"][0]


@pytest.mark.parametrize("tag", ["vol", "vol_grid", "lx", "lx_grid"])
def test_port_box_moves_golden(oracle, tag):
    """The restated volume / Lx moves against the reference's own MC_change_volume (MC.c:450-513) and MC_change_Lx
    (MC.c:515-585, including the exp(+dE/T) test of MC.c:562), driven move by move with recorded uniforms."""
    g = load_golden("box_moves")
    res, cells = g[f"{tag}_result"], g[f"{tag}_cells"]
    assert set(res.tolist()) >= ({0, 1} if tag.startswith("vol") else {-1, 0, 1})
    if tag.endswith("_grid"):
        assert (cells != g[f"{tag}_cells0"]).any()  # the sequence does cross the re-grid threshold (MC.c:481-485)
    delta, cosmax, T = g[f"{tag}_params"]
    p = oracle.PortSystem(g[f"{tag}_box"], g[f"{tag}_r"], g[f"{tag}_orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    state = {"bonds": p.total_bonds()[0], "ov": 0}
    assert -state["bonds"] == g[f"{tag}_energy0"][0]

    def vol(P, rfm, u0, u1, u2):
        acc, state["bonds"], state["ov"] = p.volume_move(P, rfm, u0, u1, u2, state["bonds"])
        return acc

    def lx(lx_max, lo, hi, u0, u1):
        acc, state["bonds"], state["ov"] = p.Lx_move(lx_max, lo, hi, u0, u1, state["bonds"])
        return acc

    replay_box_moves(g, tag, vol, lx, lambda: (p.box(), -float(state["bonds"]), state["ov"]))
    assert np.array_equal(p.state()[0], g[f"{tag}_r_final"])  # the restatement divides back like the reference: same bits


def test_port_helpers_golden(oracle):
    """Marsaglia's rejection loop + axis (utils.c:115-131) and gram_schmidt (utils.c:91-113) against what the reference's
    own functions returned for recorded inputs."""
    g = load_golden("helpers")
    L = oracle.port_lib()
    axis = np.zeros(3)
    for u, want, used in zip(g["sphere_u"], g["sphere_axis"], g["sphere_used"]):
        k = 0
        while True:  # utils.c:120-124
            ran1, ran2 = 1. - 2. * u[k], 1. - 2. * u[k + 1]
            k += 2
            if ran1 * ran1 + ran2 * ran2 < 1.:
                break
        assert k == used
        L.ppo_axis_from_pair(ran1, ran2, axis)
        assert np.array_equal(axis, want)
    for m_in, want in zip(g["gs_in"], g["gs_out"]):
        m = np.ascontiguousarray(m_in).reshape(9).copy()
        L.ppo_gram_schmidt(m)
        assert np.array_equal(m.reshape(3, 3), want)


def test_port_vs_reference_library(oracle, tmp_path):
    """The restatement against the unmodified reference, live (needs oracle/_ref built from /root/reference)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/libppref.so not built (no /root/reference here)")
    g = load_golden("gen2000_rho0.5")
    rrr = str(tmp_path / "conf.rrr")
    # write with full precision so that the reference parses back the very same doubles
    with open(rrr, "w") as f:
        f.write("0 %d %.17g %.17g %.17g\n" % (len(g["r"]), *g["box"]))
        for i in range(len(g["r"])):
            f.write("%.17g %.17g %.17g\n%.17g %.17g %.17g\n%.17g %.17g %.17g\n" % (*g["orient"][i, 0], *g["orient"][i, 1], *g["r"][i]))
    ref = oracle.RefSystem(rrr, KF_delta=0.2, KF_cosmax=0.85, Temperature=0.12)
    try:
        r, m = ref.state()
        assert np.array_equal(r, g["r"])
        port = oracle.PortSystem(ref.box(), r, m, kf_delta=0.2, kf_cosmax=0.85, temperature=0.12)
        e, ov = ref.total_energy()
        assert (-e, ov) == port.total_bonds()
        rng = np.random.default_rng(3)
        n = 3000
        idx = rng.integers(0, ref.N, n).astype(np.int32)
        disp = (rng.random((n, 3)) - 0.5) * 0.4
        u = rng.random(n)
        orient_new = np.zeros((n, 3, 3))
        for k in range(n):
            axis = rng.normal(size=3)
            axis /= np.linalg.norm(axis)
            th = rng.random() * 0.6
            orient_new[k] = ref.rotate(m[idx[k]], axis, th)
            assert np.array_equal(orient_new[k], oracle.rotate_port(m[idx[k]], axis, th))
        for commit in (False, True):
            a = ref.eval_proposals(idx, disp, orient_new, u, commit=commit)
            b = port.eval_proposals(idx, disp, orient_new, u, commit=commit)
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
        r1, m1 = ref.state()
        r2, m2 = port.state()
        assert np.array_equal(r1, r2) and np.array_equal(m1, m2)
    finally:
        ref.close()


def test_philox_known_answers(oracle):
    """Philox4x32-10 known-answer vectors of Random123 (Salmon et al., SC'11)."""
    L = oracle.port_lib()
    out = np.zeros(4, dtype=np.uint32)
    L.ppo_philox4x32(0, 0, 0, 0, 0, 0, out)
    assert [hex(x) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    L.ppo_philox4x32(0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff, out)
    assert [hex(x) for x in out] == ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    L.ppo_philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0, out)
    assert [hex(x) for x in out] == ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_rrr_reader_matches_reference_parsing(oracle, tmp_path):
    import patchyparticles_b200 as pkg
    g = load_golden("fcc")
    path = str(tmp_path / "x.rrr")
    pkg.write_rrr(path, 7, g["box"], g["r"], g["orient"])
    s1, b1, r1, m1 = pkg.read_rrr(path)
    s2, b2, r2, m2 = oracle.read_rrr(path)
    assert s1 == s2 == 7
    assert np.array_equal(b1, b2) and np.array_equal(r1, r2) and np.array_equal(m1, m2)
    # same bytes as the C writer that follows output.c:302-317
    path2 = str(tmp_path / "y.rrr")
    oracle.write_rrr(path2, 7, g["box"], g["r"], g["orient"])
    assert open(path).read() == open(path2).read()
    # header/particle-count mismatch is an error (system.c:245)
    with open(path, "a") as f:
        f.write("1 0 0\n0 1 0\n1 1 1\n")
    with pytest.raises(ValueError):
        pkg.read_rrr(path)


def test_sweep_restatement_invariants(oracle):
    """ppo_sweep: N attempts per sweep, running bond count == recount, confinement bookkeeping."""
    import patchyparticles_b200 as pkg
    box, r, m = pkg.random_fluid(1200, 0.45, seed=5)
    p = oracle.PortSystem(box, r, m, kf_delta=0.25, kf_cosmax=0.8, temperature=0.15)
    bonds, ov = p.total_bonds()
    assert ov == 0
    st = oracle.SweepStats()
    for s in range(12):
        bonds, st = p.sweep(11, s, 0.2, 0.3, bonds, st)
    assert st.tries == 12 * 1200
    assert st.tries == st.confined + st.overlapped + (st.tries - st.confined - st.overlapped)
    assert p.total_bonds() == (bonds, 0)
    assert 0.05 < st.confined / st.tries < 0.25            # 3 * E|d| / w with d uniform in +-0.1, w ~ 1.3
    raw = st.accepted / (st.tries - st.confined)
    est = st.w_accepted / st.w_tries
    assert abs(raw - est) < 0.02 and 0.2 < est < 0.9
    # a different seed gives a different trajectory, the same seed the same one
    q = oracle.PortSystem(box, r, m, kf_delta=0.25, kf_cosmax=0.8, temperature=0.15)
    b2, _ = q.total_bonds()
    st2 = oracle.SweepStats()
    for s in range(12):
        b2, st2 = q.sweep(11, s, 0.2, 0.3, b2, st2)
    assert np.array_equal(q.state()[0], p.state()[0]) and b2 == bonds


def test_checkerboard_statistics_vs_reference(oracle, tmp_path):
    """The checkerboard algorithm (serial restatement, same code path as the CUDA sweep) samples the same
    ensemble as the reference's random sequential moves: <E>/N and the acceptance estimator agree within
    3 sigma of block-averaged error bars.  N = 200, rho = 0.4, T = 0.3, cosmax = 0.92, 100000 steps each (first 20000 discarded)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    import subprocess
    subprocess.run([os.path.join(oracle.REF_DIR, "generator"), "200", "0.4"], cwd=tmp_path, check=True, stderr=subprocess.DEVNULL)
    steps, every, discard = 100000, 10, 20000
    oracle.write_input(str(tmp_path / "input"), str(tmp_path / "generated.rrr"), Temperature=0.3, Steps=steps, Disp_max=0.2, Theta_max=0.3,
                       Print_every=every, KF_cosmax=0.92, Seed=11, Log_file="none")
    subprocess.run([os.path.join(oracle.REF_DIR, "PatchyParticles"), "input"], cwd=tmp_path, check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    e = np.loadtxt(tmp_path / "energy.dat")
    a = np.loadtxt(tmp_path / "acceptance.dat")

    def blocks(x, nb=20):
        x = np.asarray(x)
        n = (len(x) // nb) * nb
        b = x[:n].reshape(nb, -1).mean(1)
        return b.mean(), b.std(ddof=1) / np.sqrt(nb)

    e_ref, se_ref = blocks(e[e[:, 0] >= discard, 1])
    a_ref, sa_ref = blocks(a[a[:, 0] >= discard, 1])
    _, box, r, m = oracle.read_rrr(str(tmp_path / "generated.rrr"))
    p = oracle.PortSystem(box, r, m, kf_delta=0.119, kf_cosmax=0.92, temperature=0.3)
    bonds, _ = p.total_bonds()
    st, E, A, last = oracle.SweepStats(), [], [], (0, 0)
    for s in range(steps):
        bonds, st = p.sweep(5, s, 0.2, 0.3, bonds, st)
        if s % every == 0 and s >= discard:
            E.append(-bonds / len(r))
            A.append((st.w_accepted - last[0]) / max(st.w_tries - last[1], 1))
        if s % every == 0:
            last = (st.w_accepted, st.w_tries)
    e_port, se_port = blocks(E)
    a_port, sa_port = blocks(A)
    print(f"E/N ref {e_ref:.4f} +- {se_ref:.4f} port {e_port:.4f} +- {se_port:.4f}; acc ref {a_ref:.4f} +- {sa_ref:.4f} port {a_port:.4f} +- {sa_port:.4f}")
    assert -0.4 < e_ref < -0.2
    assert abs(e_port - e_ref) < 3 * np.hypot(se_ref, se_port)
    assert abs(a_port - a_ref) < 3 * np.hypot(sa_ref, sa_port)


def test_abi_exports_every_declared_symbol():
    """libpatchy_b200.so loads and exports every function include/pp_b200.h declares (no compute calls)."""
    import patchyparticles_b200 as pkg
    header = open(os.path.join(ROOT, "include", "pp_b200.h")).read()
    declared = set(re.findall(r"\b(pp_[a-zA-Z_0-9]+)\s*\(", header))
    declared -= {"pp_handle", "pp_params", "pp_observables"}
    assert declared == set(pkg.engine.ABI_SYMBOLS), declared ^ set(pkg.engine.ABI_SYMBOLS)
    lib = ctypes.CDLL(pkg.library_path())
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert lib.pp_abi_version() == 4
    assert ctypes.sizeof(pkg.Params) >= 120


def test_no_cpu_fallback():
    """Without a CUDA device the engine refuses to start instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import patchyparticles_b200 as pkg
    with pytest.raises(Exception) as e:
        pkg.Engine()
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under patchyparticles_b200/ or host/ may reference it."""
    bad = []
    for base in ("patchyparticles_b200", "host", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".c", ".h", "Makefile")):
                    text = open(os.path.join(dirpath, fn), errors="ignore").read()
                    if re.search(r'#include\s+"[^"]*oracle|import\s+_oracle|from\s+_oracle|libpporacle|libppref|ppo_[a-z_]+\s*\(', text):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad


def test_gc_golden_fixture(oracle, tmp_path):
    """tests/golden/gc_moves.npz (the reference's own MC_add_remove, MC.c:342-448, fed recorded uniforms): the file is
    consistent with the move it records, and -- where the reference library is present -- the reference replayed live takes
    the same decisions, so the fixture is what tools/make_golden.py gc_moves produces."""
    g = load_golden("gc_moves")
    kind, acc, used, Ns, en = g["kind"], g["accepted"], g["used"], g["N"], g["energies"]
    assert set(np.unique(kind)) <= {2, 3} and (acc[kind == 2].sum(), acc[kind == 3].sum()) == (29, 20)
    n_prev = len(g["r"])
    for k in range(len(kind)):
        assert Ns[k] == n_prev + (acc[k] if kind[k] == 2 else -acc[k])
        assert (g["u"][k, 0] < 0.5) == (kind[k] == 2)          # MC.c:344
        if kind[k] == 3:
            assert used[k] == 3                                 # which move, which particle, acceptance (MC.c:344, 383, 387)
        else:
            assert used[k] >= 10 and (used[k] - 4) % 2 in (0, 1)  # 1 + 3 + three Marsaglia points (>= 2 each) [+ acceptance without an overlap]
        n_prev = Ns[k]
    assert np.all(en == np.round(en)) and len(g["r_final"]) == Ns[-1]
    if not oracle.have_ref():
        return
    rrr = str(tmp_path / "warm.rrr")
    oracle.write_rrr(rrr, 0, g["box"], g["r"], g["orient"].reshape(-1, 9))
    delta, cosmax, T, z, n_min, n_max = g["params"]
    ref = oracle.RefSystem(rrr, Ensemble=1, Activity=z, GC_N_max=int(n_max), KF_delta=delta, KF_cosmax=cosmax, Temperature=T)
    e, ov = ref.total_energy()
    assert (e, ov) == (g["energy0"][0], 0)
    ref.set_running_energy(e)
    for k in range(120):
        assert ref.add_remove(g["u"][k]) == (kind[k], acc[k], used[k], Ns[k], en[k]), f"move {k}"
    ref.close()


def test_neighbour_queries_fixture(oracle, tmp_path):
    """tests/golden/neighbour_queries.npz: symmetric (q lists p with the patches swapped), consistent with the bond counts of
    warm1000.npz at the wide state point, every bonded pair inside the range lists; regenerated live where the reference is."""
    g = load_golden("neighbour_queries")
    w = load_golden("warm1000")
    assert np.array_equal(g["wide_count"], w["bond_counts"])
    for tag in ("narrow", "wide"):
        cnt, q, onp, onq = g[f"{tag}_count"], g[f"{tag}_q"], g[f"{tag}_onp"], g[f"{tag}_onq"]
        for i in range(len(cnt)):
            for k in range(cnt[i]):
                j = int(q[i, k])
                assert i in q[j, :cnt[j]].tolist() and j in g["range_nb"][i, :g["range_count"][i]].tolist()
                if tag == "narrow":  # one neighbour per patch: the pair is reported from the other side with the patches swapped
                    kk = q[j, :cnt[j]].tolist().index(i)
                    assert (onp[j, kk], onq[j, kk]) == (onq[i, k], onp[i, k])
    if not oracle.have_ref():
        return
    rrr = str(tmp_path / "warm.rrr")
    oracle.write_rrr(rrr, 0, g["box"], g["r"], g["orient"].reshape(-1, 9))
    delta, T = g["params"]
    ref = oracle.RefSystem(rrr, KF_delta=delta, KF_cosmax=0.8, Temperature=T)
    c, q, onp, onq = ref.patch_pairs(8)
    ref.close()
    assert np.array_equal(c, g["wide_count"]) and np.array_equal(q, g["wide_q"]) and np.array_equal(onp, g["wide_onp"])


def test_port_gc_moves_golden(oracle):
    """The restated MC_add_remove against the reference's own (tests/golden/gc_moves.npz): every decision, the uniforms consumed,
    N and the running energy after each of the 400 moves, and the final configuration index for index, bit for bit."""
    g = load_golden("gc_moves")
    delta, cosmax, T, z, n_min, n_max = g["params"]
    p = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
    bonds, ov = p.total_bonds()
    assert (-bonds, ov) == (g["energy0"][0], 0)
    for k in range(len(g["u"])):
        kind, acc, used, N, bonds = p.add_remove(z, int(n_min), int(n_max), g["u"][k], bonds)
        assert (kind, acc, used, N, -bonds) == (g["kind"][k], g["accepted"][k], g["used"][k], g["N"][k], g["energies"][k]), f"move {k}"
    r, m = p.state()
    assert np.array_equal(r, g["r_final"]) and np.array_equal(m, g["orient_final"])
    assert p.total_bonds() == (bonds, 0)


def test_port_neighbour_queries_golden(oracle):
    """The restated neighbour queries (avb.c:43-75, vmmc.c:178) against the reference's MC_interact over all pairs."""
    g = load_golden("neighbour_queries")
    delta, T = g["params"]
    N = len(g["r"])
    for tag, cosmax in (("narrow", 0.98), ("wide", 0.8)):
        p = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=cosmax, temperature=T)
        cnt, q, onp = g[f"{tag}_count"], g[f"{tag}_q"], g[f"{tag}_onp"]
        for i in range(N):
            nb, num = p.patch_neighbours(i)
            assert num == cnt[i]
            for patch in range(4):
                on_patch = [int(q[i, k]) for k in range(cnt[i]) if onp[i, k] == patch]
                assert (nb[patch] == -1) if not on_patch else (nb[patch] in on_patch)
    p = oracle.PortSystem(g["box"], g["r"], g["orient"], kf_delta=delta, kf_cosmax=0.98, temperature=T)
    for i in range(0, N, 7):
        n, nb = p.range_neighbours(i, 24)
        assert n == g["range_count"][i] and sorted(nb[:n].tolist()) == sorted(g["range_nb"][i, :n].tolist())
