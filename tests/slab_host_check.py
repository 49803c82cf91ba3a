"""Run under torchrun with the gloo backend (CPU): the host-side plumbing of the slab decomposition --
ring neighbours, count exchange, variable-size record exchange (with two ranks both neighbours are the
same peer), and the layer partition."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from patchyparticles_b200.slab import exchange_counts, exchange_with_neighbours, layer_partition, neighbours  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    left, right = neighbours(rank, world)
    cpu = torch.device("cpu")
    # counts: rank r announces 100 + r to its left and 200 + r to its right
    from_l, from_r = exchange_counts(rank, world, 100 + rank, 200 + rank, cpu)
    assert from_l == 200 + left, (rank, from_l)     # what my left neighbour sent to ITS right
    assert from_r == 100 + right, (rank, from_r)    # what my right neighbour sent to ITS left
    # records of different lengths in both directions, several tensors per direction (as the halo refresh does)
    n_l, n_r = 3 + rank, 5 + 2 * rank
    to_left = [torch.full((13 * n_l,), float(10 * rank + 1)), torch.full((n_l,), float(10 * rank + 2))]
    to_right = [torch.full((13 * n_r,), float(10 * rank + 3)), torch.full((n_r,), float(10 * rank + 4))]
    nl_in, nr_in = exchange_counts(rank, world, n_l, n_r, cpu)
    assert nl_in == 5 + 2 * left and nr_in == 3 + right
    from_left = [torch.zeros(13 * nl_in), torch.zeros(nl_in)]
    from_right = [torch.zeros(13 * nr_in), torch.zeros(nr_in)]
    exchange_with_neighbours(rank, world, to_left, to_right, from_left, from_right)
    assert torch.all(from_left[0] == 10 * left + 3) and torch.all(from_left[1] == 10 * left + 4)
    assert torch.all(from_right[0] == 10 * right + 1) and torch.all(from_right[1] == 10 * right + 2)
    # one-directional exchange with empty lists on the other side (halo refresh after a pass block)
    a, b = torch.full((7,), float(rank)), torch.zeros(7)
    exchange_with_neighbours(rank, world, [a], [], [], [b])
    assert torch.all(b == right)
    # zero-length messages are skipped consistently on both ends
    exchange_with_neighbours(rank, world, [torch.zeros(0)], [torch.zeros(0)], [torch.zeros(0)], [torch.zeros(0)])
    # the partition is the same on every rank, covers all layers with even boundaries
    parts = layer_partition(170, world)
    assert parts[0][0] == 0 and parts[-1][1] == 170
    assert all(p[0] % 2 == 0 and p[1] % 2 == 0 and p[1] - p[0] >= 2 for p in parts)
    assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
