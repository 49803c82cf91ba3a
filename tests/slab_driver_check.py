"""Run under torchrun with the gloo backend (CPU): the sweep loop of SlabEngine (patchyparticles_b200/slab.py) driven
against a stand-in engine that keeps its "device" arrays in host memory.

nccl transport: the stand-in only moves data when the driver tells it to, so after a few sweeps every halo layer must
hold exactly what its owner last wrote -- direction, order and sizes of the halo refresh and of the migration
exchange are checked with real messages between the ranks.
peer transport: the data path is inside the library (mirror stores from the sweep kernels, phase numbers between
neighbours); the driver's job is the call sequence, which is compared with the one include/pp_b200.h prescribes
(pp_slab_migrate at every grid shift, pp_slab_sweeps never crossing a shift)."""
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from patchyparticles_b200.slab import RECORD_DOUBLES, SlabEngine, neighbours  # noqa: E402

PERIOD = 4


def n_first(r):
    return 3 + r


def n_last(r):
    return 4 + 2 * r


class StandInEngine:
    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.left, self.right = neighbours(rank, world)
        self.step = 0
        self.log = []
        self.layer = {}  # name -> (pos, ori, posf)
        for name, n in (("first", n_first(rank)), ("last", n_last(rank)), ("halo_l", n_last(self.left)), ("halo_r", n_first(self.right))):
            self.layer[name] = (np.zeros(4 * n), np.zeros(9 * n), np.zeros(4 * n, dtype=np.int32))
        self.outbox, self.inbox = [None, None], [None, None]
        self.rebuilds = 0

    # ---- what the driver calls ---------------------------------------------------------------------
    def slab_setup(self, rank, world):
        assert (rank, world) == (self.rank, self.world)

    def sync(self):
        pass

    def slab_range(self):
        return 0, 0, 0, 0

    def _stamp(self, name, tag):
        v = 1000.0 * self.step + 10.0 * self.rank + tag
        for a in self.layer[name]:
            a[:] = v

    def slab_sweep_begin(self):
        self.log.append(("begin", self.step))
        return self.step % PERIOD == 0

    def slab_outbox(self, side):
        n = (2 + self.rank) if side == 0 else (1 + 3 * self.rank)
        self.outbox[side] = np.full(RECORD_DOUBLES * n, 100.0 * self.step + 10.0 * self.rank + side)
        return self.outbox[side].ctypes.data, n

    def slab_inbox(self, side, count):
        self.inbox[side] = np.zeros(RECORD_DOUBLES * count)
        return self.inbox[side].ctypes.data

    def slab_sweep_rebuild(self):
        # from the left neighbour comes what IT sent to its right (side 1), and vice versa
        assert len(self.inbox[0]) == RECORD_DOUBLES * (1 + 3 * self.left) and np.all(self.inbox[0] == 100.0 * self.step + 10.0 * self.left + 1)
        assert len(self.inbox[1]) == RECORD_DOUBLES * (2 + self.right) and np.all(self.inbox[1] == 100.0 * self.step + 10.0 * self.right + 0)
        self.rebuilds += 1
        self.log.append(("rebuild", self.step))

    def slab_pass_block(self, cx):
        self.log.append(("pass", self.step, cx))
        self._stamp("first" if cx == 0 else "last", 1 + cx)  # block 0 moves the first owned layer, block 1 the last one

    def slab_halo(self, side, recv):
        name = (("halo_l", "halo_r") if recv else ("first", "last"))[side]
        pos, ori, posf = self.layer[name]
        return pos.ctypes.data, ori.ctypes.data, posf.ctypes.data, len(posf) // 4

    def slab_sweep_end(self):
        self.log.append(("end", self.step))
        self.step += 1

    # peer transport
    def slab_migrate(self):
        self.log.append(("migrate?", self.step))
        if self.step % PERIOD != 0 or getattr(self, "_migrated_at", -1) == self.step:
            return False
        self._migrated_at = self.step
        self.rebuilds += 1
        return True

    def slab_sweeps(self, max_sweeps):
        """like pp_slab_sweeps: refuses to start a shift period that has not been migrated, stops before the next one"""
        done = 0
        while done < max_sweeps:
            if self.step % PERIOD == 0 and getattr(self, "_migrated_at", -1) != self.step:
                assert done > 0, "pp_slab_sweeps at the start of a shift period without pp_slab_migrate"
                break
            for cx in (0, 1):
                self.slab_pass_block(cx)
            self.slab_sweep_end()
            done += 1
        self.log.append(("sweeps", max_sweeps, done))
        return done


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo")
    left, right = neighbours(rank, world)

    # ---- send/recv transport: real messages, real data ------------------------------------------------
    eng = StandInEngine(rank, world)
    slab = SlabEngine(rank, world, 0, engine=eng, transport="nccl")
    slab.sweep(11)
    assert eng.step == 11 and eng.rebuilds == 3
    # last refresh happened in sweep 10: block 0 stamped the first layers (tag 1), block 1 the last layers (tag 2)
    for a in eng.layer["halo_r"]:
        assert len(a) and np.all(a == type(a[0])(1000.0 * 10 + 10.0 * right + 1)), (rank, a[:3])
    for a in eng.layer["halo_l"]:
        assert np.all(a == type(a[0])(1000.0 * 10 + 10.0 * left + 2)), (rank, a[:3])
    order = [e for e in eng.log if e[0] in ("pass", "end", "rebuild")]
    assert order[:4] == [("rebuild", 0), ("pass", 0, 0), ("pass", 0, 1), ("end", 0)]
    assert [e[1] for e in eng.log if e[0] == "rebuild"] == [0, 4, 8]

    # ---- peer transport: the call sequence ------------------------------------------------------------
    eng = StandInEngine(rank, world)
    slab = SlabEngine(rank, world, 0, engine=eng, transport="peer")
    slab.sweep(3)
    slab.sweep(8)   # crosses two grid shifts, starts in the middle of a period
    assert eng.step == 11 and eng.rebuilds == 3
    calls = [e for e in eng.log if e[0] in ("pass", "end")]
    expect = []
    for s in range(11):
        expect += [("pass", s, 0), ("pass", s, 1), ("end", s)]
    assert calls == expect
    # a grid shift is never skipped: every sweep that starts a period is preceded by a successful migrate at that step
    for s in (0, 4, 8):
        i_m = eng.log.index(("migrate?", s))
        assert i_m < eng.log.index(("pass", s, 0))
    # the library-side loop never ran across a shift
    assert all(e[2] <= PERIOD for e in eng.log if e[0] == "sweeps")
    dist.barrier()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
