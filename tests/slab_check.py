"""Launched under torchrun by tests/test_gpu_slab.py (one process per GPU): a slab-decomposed run must
follow the single-GPU trajectory bit for bit (same seed, same global cell streams), keep the global
energy invariant (running == recount) and lose no particle across migrations."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import patchyparticles_b200 as pkg  # noqa: E402
from patchyparticles_b200.slab import SlabEngine  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, rho, sweeps = int(sys.argv[1]), float(sys.argv[2]), int(sys.argv[3])
    params = dict(kf_delta=0.25, kf_cosmax=0.8, temperature=0.15, disp_max=0.2, theta_max=0.3, seed=31, shift_every=4)
    box, r, m = pkg.random_fluid(N, rho, seed=9)

    slab = SlabEngine(rank, world, local, **params)
    slab.upload(box, r, m)
    b0, ov0 = slab.total_bonds()
    x0, x1, nx, n_local = slab.eng.slab_range()
    print(f"[rank {rank}] layers [{x0}, {x1}) of {nx}, {n_local} local particles, bonds {b0}", flush=True)
    slab.sweep(sweeps)
    obs = slab.observables()
    R, M = slab.download()
    b_chk, ov_chk, ok = slab.check_energy()  # (re-orthonormalises the orientations: after the download)

    failures = []
    if rank == 0:
        ref = pkg.Engine(device=local, **params)
        ref.upload(box, r, m)
        rb0, rov0 = ref.total_bonds()
        ref.sweep(sweeps)
        o = ref.observables()
        rr, rm = ref.download()
        if (b0, ov0) != (rb0, rov0):
            failures.append(f"initial bonds {b0, ov0} != {rb0, rov0}")
        if obs["energy"] != o.energy:
            failures.append(f"energy {obs['energy']} != {o.energy}")
        for k, v in (("tries", o.tries[0]), ("accepted", o.accepted[0]), ("confined", o.confined), ("overlapped", o.overlapped),
                     ("pair_checks", o.pair_checks)):
            if obs[k] != v:
                failures.append(f"{k} {obs[k]} != {v}")
        if not np.array_equal(R, rr):
            failures.append(f"positions differ for {(R != rr).any(axis=1).sum()} particles")
        if not np.array_equal(M, rm):
            failures.append(f"orientations differ for {(M != rm).any(axis=(1, 2)).sum()} particles")
        if not ok or -b_chk != obs["energy"]:
            failures.append(f"energy check failed: recount {b_chk}, running {obs['energy']}, overlap {ov_chk}")
        moved = np.abs(R - r).max()
        print(f"[rank 0] {world} slabs vs 1 GPU after {sweeps} sweeps: E = {obs['energy']}, accepted {obs['accepted']}/{obs['tries']}, "
              f"max displacement {moved:.3f}, failures: {failures}", flush=True)
        ref.close()
    flag = torch.tensor([len(failures)], device=torch.device("cuda", local))
    dist.broadcast(flag, 0)
    slab.close()
    dist.destroy_process_group()
    sys.exit(1 if int(flag.item()) else 0)


if __name__ == "__main__":
    main()
