"""The x-slab decomposition on ONE GPU: a pp_group of n "virtual slabs" runs through the same code as a multi-GPU run --
mirror stores of accepted boundary moves into the neighbour's halo, migration records through the neighbour's
mailboxes, phase ordering between the engines, box moves decided on the sums -- and has to follow the single engine's
trajectory bit for bit (same seed, random streams keyed by the global cell).  With >= 2 GPUs the same group is also
spread over the devices (peer access), and the multi-process transport is covered by tests/test_gpu_slab.py."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, load_golden, replay_box_moves

pytestmark = pytest.mark.gpu

PARAMS = dict(kf_delta=0.25, kf_cosmax=0.8, temperature=0.15, disp_max=0.2, theta_max=0.3, seed=31, shift_every=4)


@pytest.fixture(scope="module")
def pp():
    import patchyparticles_b200 as pkg
    pkg.load_library()
    return pkg


def n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def compare_with_single(pp, group, params, box, r, m, sweeps, chunks=(None,)):
    ref = pp.Engine(**params)
    ref.upload(box, r, m)
    group.upload(box, r, m)
    assert group.total_bonds() == ref.total_bonds()
    done = 0
    for c in chunks:  # the group is driven in pieces that start in the middle of a shift period too
        k = sweeps - done if c is None else c
        group.sweep(k)
        ref.sweep(k)
        done += k
    o, q = ref.observables(), group.observables()
    assert q.energy == o.energy and q.step == o.step == sweeps
    for name in ("confined", "overlapped", "pair_checks", "w_tries", "w_accepted"):
        assert getattr(q, name) == getattr(o, name), name
    assert list(q.tries) == list(o.tries) and list(q.accepted) == list(o.accepted)
    assert np.array_equal(np.array(q.box), np.array(o.box)) and q.volume == o.volume
    R, M = group.download()
    rr, mm = ref.download()
    assert np.array_equal(R, rr), f"positions differ for {(R != rr).any(axis=1).sum()} particles"
    assert np.array_equal(M, mm)
    b, ov, ok = group.check_energy()
    assert ok and ov == 0 and -b == q.energy
    assert ref.check_energy() == (b, ov, ok)
    assert 0 < np.abs(R - r).max()
    ref.close()
    return o


@pytest.mark.parametrize("n,N,rho,sweeps,chunks", [(2, 20000, 0.5, 9, (None,)), (3, 20000, 0.5, 9, (2, 5, None)), (4, 40000, 0.55, 9, (None,)),
                                                  (2, 3000, 0.3, 13, (1, 1, None)), (5, 60000, 0.6, 6, (None,))])
def test_virtual_slabs_match_single_engine(pp, n, N, rho, sweeps, chunks):
    box, r, m = pp.random_fluid(N, rho, seed=9)
    g = pp.Group(n, **PARAMS)
    o = compare_with_single(pp, g, PARAMS, box, r, m, sweeps, chunks)
    assert 0 < o.accepted[0] < o.tries[0] == sweeps * N
    ranges = g.slab_ranges()
    assert ranges[0][0] == 0 and ranges[-1][1] == ranges[0][2] and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    assert sum(x[3] for x in ranges) > N  # owned + halo copies
    g.close()


@pytest.mark.parametrize("n,N,rho", [(3, 30000, 0.5), (4, 40000, 0.55)])
def test_balanced_slabs(pp, n, N, rho):
    """x_layers_multiple = 2 n: the number of x layers is a multiple of 2 n, every slab owns the same number of them, and a single
    engine on the same grid follows the same trajectory bit for bit (the grid, not the decomposition, decides the trajectory)."""
    box, r, m = pp.random_fluid(N, rho, seed=9)
    params = dict(PARAMS, x_layers_multiple=2 * n)
    plain = pp.Engine(**PARAMS)
    plain.upload(box, r, m)
    nx_plain = plain.observables().n_side[0]
    plain.close()
    g = pp.Group(n, **params)
    o = compare_with_single(pp, g, params, box, r, m, 6)
    ranges = g.slab_ranges()
    nx = ranges[0][2]
    assert nx % (2 * n) == 0 and nx == nx_plain - nx_plain % (2 * n) and o.n_side[0] == nx
    assert len({x[1] - x[0] for x in ranges}) == 1  # the same number of layers everywhere
    g.close()


@pytest.mark.parametrize("kernel", [dict(lanes_per_cell=1), dict(cells_max_occupancy=1, cells_mid_occupancy=2), dict(cells_max_occupancy=2, cell_lanes=2),
                                    dict(lanes_per_cell=8)])
def test_virtual_slabs_every_kernel(pp, kernel):
    """the mirror stores sit in every kernel that commits a move: fused cells kernel, its stand-alone configurations, the
    serial fallback inside them, the tile kernel"""
    box, r, m = pp.random_fluid(30000, 0.6, seed=4)
    params = dict(PARAMS, **kernel)
    g = pp.Group(3, **params)
    compare_with_single(pp, g, params, box, r, m, 6)
    g.close()


def test_virtual_slabs_large(pp):
    """passes large enough for the fused launch's striding, crowded cells on 8 lanes, the rest on one thread per cell"""
    box, r, m = pp.random_fluid(400000, 0.6, seed=2)
    params = dict(PARAMS, kf_delta=0.119, kf_cosmax=0.92, disp_max=0.1)
    g = pp.Group(2, **params)
    compare_with_single(pp, g, params, box, r, m, 5)
    g.close()


@pytest.mark.parametrize("tag", ["vol", "vol_grid", "lx", "lx_grid"])
@pytest.mark.parametrize("n", [2, 3])
def test_group_box_moves_golden(pp, tag, n):
    """Box moves under the decomposition against the reference's own MC_change_volume / MC_change_Lx (golden vectors):
    every engine rescales and counts its layers, the decision is taken on the sums; the sequences that cross the re-grid
    threshold (MC.c:481-485) re-partition the slabs."""
    g = load_golden("box_moves")
    delta, cosmax, T = g[f"{tag}_params"]
    P, rfm, lx_max, lyz_min, lyz_max = g[f"{tag}_keys"]
    common = dict(kf_delta=float(delta), kf_cosmax=float(cosmax), temperature=float(T))
    if tag.startswith("vol"):
        grp = pp.Group(n, ensemble=5, P=float(P), rescale_factor_max=float(rfm), **common)
    else:
        grp = pp.Group(n, Lx_move=1, Lx_change_max=float(lx_max), Lyz_min=float(lyz_min), Lyz_max=float(lyz_max), **common)
    grp.upload(g[f"{tag}_box"], g[f"{tag}_r"], g[f"{tag}_orient"])
    assert grp.total_bonds() == (-g[f"{tag}_energy0"][0], 0)

    def observe():
        o = grp.observables()
        return np.array(o.box), o.energy, None

    replay_box_moves(g, tag, lambda P_, rfm_, a, b, c: grp.volume_move(a, b, c), lambda mx, lo, hi, a, b: grp.Lx_move(a, b), observe)
    o = grp.observables()
    slot = 6 if tag.startswith("vol") else 7
    assert o.tries[slot] == len(g[f"{tag}_u"]) and o.accepted[slot] == int((g[f"{tag}_result"] == 1).sum())
    r, _ = grp.download()
    assert np.abs(r - g[f"{tag}_r_final"]).max() < 1e-12
    b, ov, ok = grp.check_energy()
    assert ok and -b == g[f"{tag}_energies"][-1]
    grp.sweep(5)
    assert grp.check_energy()[2]
    grp.close()


@pytest.mark.parametrize("mode", ["npt", "lx"])
def test_group_ensemble_run_matches_single_engine(pp, mode):
    """whole do_NPT / do_NVT + Lx_move steps (MC.c:19-29, 60-68): sweeps interleaved with box moves from the shared stream,
    some accepted, bit-identical to the single engine (box, volume, energy, state)"""
    box, r, m = pp.random_fluid(4000, 0.35, seed=12)
    if mode == "npt":
        params = dict(PARAMS, ensemble=5, P=0.05, rescale_factor_max=4e-3, temperature=0.25)
    else:
        L = float(box[1])
        params = dict(PARAMS, Lx_move=1, Lx_change_max=0.05, Lyz_min=L - 0.2, Lyz_max=L + 0.2, temperature=0.25)
    g = pp.Group(2, **params)
    o = compare_with_single(pp, g, params, box, r, m, 60, (7, None))
    slot = 6 if mode == "npt" else 7
    assert o.tries[slot] > 20 and 0 < o.accepted[slot] < o.tries[slot]
    g.close()


def test_group_over_devices(pp):
    """the same group spread over the GPUs of the box (peer access), including one device used twice"""
    k = n_gpus()
    if k < 2:
        pytest.skip("needs at least 2 GPUs")
    box, r, m = pp.random_fluid(60000, 0.55, seed=3)
    devices = list(range(min(k, 4)))
    g = pp.Group(len(devices) + 1, devices=devices + [0], **PARAMS)
    compare_with_single(pp, g, PARAMS, box, r, m, 9)
    g.close()


def test_c_host_with_gpus_key(pp, tmp_path):
    """host/patchy_b200 with GPUs = 3 (three slabs on device 0) writes the same last.rrr and energy.dat as with one engine"""
    host = os.path.join(ROOT, "host", "patchy_b200")
    box, r, m = pp.random_fluid(6000, 0.45, seed=8)
    conf = str(tmp_path / "init.rrr")
    pp.write_rrr(conf, 0, box, r, m)
    outs = []
    for gpus in (1, 3):
        d = tmp_path / f"run{gpus}"
        d.mkdir()
        with open(d / "input", "w") as f:
            f.write(f"Dynamics = 0\nEnsemble = 0\nTemperature = 0.15\nSteps = 24\nDisp_max = 0.2\nTheta_max = 0.3\n"
                    f"Initial_conditions_file = {conf}\nPrint_every = 8\nSave_every = 1000\nConfiguration_folder = .\n"
                    f"KF_delta = 0.25\nKF_cosmax = 0.8\nSeed = 5\nGPUs = {gpus}\nGPU_devices = {' '.join(['0'] * gpus)}\n")
        p = subprocess.run([host, "input"], cwd=d, capture_output=True, text=True, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append((open(d / "last.rrr").read(), open(d / "energy.dat").read(), open(d / "acceptance.dat").read()))
    assert outs[0] == outs[1]


def test_upload_select_from_pinned_and_pageable_memory(pp):
    """pp_upload_select: a slab engine handed the whole configuration keeps the particles of its layers -- selected and
    gathered by a kernel reading pinned host memory, or on the host from pageable memory -- exactly the set pp_slab_select
    names, whatever the order they arrive in."""
    import torch
    box, r, m = pp.random_fluid(50000, 0.5, seed=6)
    m9 = m.reshape(-1, 9)
    r_pin, m_pin = torch.from_numpy(r.copy()).pin_memory().numpy(), torch.from_numpy(m9.copy()).pin_memory().numpy()
    for rank, world in ((0, 2), (1, 3), (4, 5)):
        results = []
        for how in ("subset", "pinned", "pageable", "pinned again"):
            e = pp.Engine(**PARAMS) if how != "pinned again" else results[1][2]
            if how != "pinned again":
                e.slab_setup(rank, world)
            if how == "subset":
                mine = e.slab_select(rank, world, box, r)
                ids = np.flatnonzero(mine).astype(np.int32)
                e.upload_subset(len(r), ids, box, r[ids], m9[ids])
            elif how == "pageable":
                e.upload_select(box, r, m9)
            else:
                e.upload_select(box, r_pin, m_pin)
            ids_own, r_own, m_own = e.slab_download()
            order = np.argsort(ids_own)
            results.append(((e.slab_range(), e.total_bonds()), (ids_own[order], r_own[order], m_own[order]), e))
        for res in results[1:]:
            assert res[0] == results[0][0]
            for a, b in zip(res[1], results[0][1]):
                assert np.array_equal(a, b)
        assert np.array_equal(results[0][1][1], r[results[0][1][0]])
        for res in results[:3]:
            res[2].close()
