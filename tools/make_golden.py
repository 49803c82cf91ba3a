#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libppref.so, built by
oracle/Makefile from /root/reference).  Run in the build container only:

    make -C oracle && python tools/make_golden.py

Each file holds the configuration as the reference parsed it (system.c:213-245) and the
reference's own answers: total energy, per-particle bond counts, and the (deltaE, overlap, accept)
decisions of MC.c:328-335 on seeded proposal lists, frozen and sequential.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"


def proposals(rng, ref, m, n, disp_max, theta_max):
    N = ref.N
    idx = rng.integers(0, N, n).astype(np.int32)
    disp = (rng.random((n, 3)) - 0.5) * disp_max
    u = rng.random(n)
    orient_new = np.zeros((n, 3, 3))
    axis = np.zeros(3)
    for k in range(n):
        while True:
            a, b = 1 - 2 * rng.random(), 1 - 2 * rng.random()
            if a * a + b * b < 1:
                break
        O.port_lib().ppo_axis_from_pair(a, b, axis)
        orient_new[k] = ref.rotate(m[idx[k]], axis, rng.random() * theta_max)  # utils.c:204-209
    return idx, disp, orient_new, u


def one(name, rrr, variants, n_prop=0, disp_max=0.3, theta_max=0.5, T=0.1, seed=7):
    out = {}
    for tag, keys in variants.items():
        ref = O.RefSystem(rrr, Temperature=T, **keys)
        r, m = ref.state()
        out["box"], out["r"], out["orient"] = ref.box(), r, m
        e, ov = ref.total_energy()
        counts, _ = ref.bond_counts()
        out[f"energy_{tag}"] = np.array([e, ov])
        out[f"bond_counts_{tag}"] = counts
        out[f"params_{tag}"] = np.array([keys.get("KF_delta", 0.119), keys.get("KF_cosmax", 0.98), T])
        out[f"ref_cells_{tag}"] = ref.cells()
        if n_prop and tag == "a":
            rng = np.random.default_rng(seed)
            idx, disp, on, u = proposals(rng, ref, m, n_prop, disp_max, theta_max)
            dE, ovl, acc = ref.eval_proposals(idx, disp, on, u, commit=False)
            out.update(prop_idx=idx, prop_disp=disp, prop_orient=on, prop_u=u, prop_dE=dE, prop_overlap=ovl,
                       prop_accept=acc)
            n_seq = min(n_prop, 1024)
            dE, ovl, acc = ref.eval_proposals(idx[:n_seq], disp[:n_seq], on[:n_seq], u[:n_seq], commit=True)
            r2, m2 = ref.state()
            e2, _ = ref.total_energy()
            out.update(seq_dE=dE, seq_overlap=ovl, seq_accept=acc, seq_r=r2, seq_orient=m2, seq_energy=np.array([e2]))
        ref.close()
        print(name, tag, "E =", e, "overlap =", ov, "bond hist", np.bincount(counts))
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)


def bond_lists():
    """tests/golden/bond_lists.npz: bonded neighbours of every particle by the reference's MC_interact over all
    pairs (what output_print_bonds, output.c:213-268, is meant to list), rows ascending, -1 padded."""
    out = {}
    tmp = "/tmp/_golden_gen.rrr"
    O.run_generator(2000, 0.5, tmp)
    for name, rrr, keys in (("fcc", os.path.join(REF, "fcc.rrr"), {"KF_cosmax": 0.92}),
                            ("gen2000_rho0.5", tmp, {"KF_cosmax": 0.8, "KF_delta": 0.25})):
        ref = O.RefSystem(rrr, Temperature=0.1, **keys)
        counts, nb = ref.bond_lists(12)
        assert counts.max() <= 12
        out[f"{name}_counts"], out[f"{name}_neighbours"] = counts, O.sorted_bond_lists(nb)
        out[f"{name}_params"] = np.array([keys.get("KF_delta", 0.119), keys.get("KF_cosmax", 0.98)])
        print("bond lists", name, "bonds per particle", np.bincount(counts))
        ref.close()
    np.savez_compressed(os.path.join(GOLD, "bond_lists.npz"), **out)


def main():
    os.makedirs(GOLD, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "bond_lists":
        return bond_lists()
    one("fcc", os.path.join(REF, "fcc.rrr"), {"a": {}, "b": {"KF_cosmax": 0.92}}, n_prop=2048, T=0.11)
    one("bcc", os.path.join(REF, "bcc.rrr"), {"a": {}, "b": {"KF_delta": 0.135}}, n_prop=1024)
    tmp = "/tmp/_golden_gen.rrr"
    O.run_generator(1000, 0.1, tmp)
    one("gen1000_rho0.1", tmp, {"a": {}}, n_prop=2048)
    O.run_generator(2000, 0.5, tmp)
    one("gen2000_rho0.5", tmp, {"a": {}, "b": {"KF_cosmax": 0.8, "KF_delta": 0.25}}, n_prop=4096)
    # known answers for the large generator outputs (configuration regenerated at test time by the
    # reference's deterministic generator when oracle/_ref/ is present)
    bond_lists()
    known = {}
    for N, rho in [(32768, 0.3), (65536, 0.5)]:
        O.run_generator(N, rho, tmp)
        ref = O.RefSystem(tmp)
        e, ov = ref.total_energy()
        r, _ = ref.state()
        known[f"{N}_{rho}"] = np.array([e, ov, float(np.sum(r))])
        print("generator", N, rho, "E =", e)
        ref.close()
    np.savez(os.path.join(GOLD, "generator_known.npz"), **known)


if __name__ == "__main__":
    main()
