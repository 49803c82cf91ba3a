"""x-slab decomposition of one simulation across the GPUs of a box: one process (and one
:class:`Engine`) per GPU, ``torch.distributed`` for the plumbing (NCCL over NVLink for the device
buffers, gloo in the CPU tests of this module's host logic).

The reference has no decomposition (single thread, main.c:26-143).  Here every engine owns whole
pairs of x layers of the global cell grid plus one halo layer per side.  The
colour passes are ordered with the x parity slowest, so a sweep needs exactly two halo refreshes:
after the four even-x passes every engine's first owned layer goes to its left neighbour, after the
four odd-x passes its last owned layer goes to its right neighbour.  When the random grid shift is
redrawn (every ``shift_every`` sweeps) particles that crossed a slab boundary migrate, together with
the fresh halo layers, in one neighbour exchange of 13-double records.

Random streams are keyed by the global cell index, so a decomposed run follows the single-GPU
trajectory bit for bit (tests/test_gpu_slab.py).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

from .engine import Engine

RECORD_DOUBLES = 13
_DEBUG = bool(os.environ.get("PP_SLAB_DEBUG"))
PARANOID = bool(os.environ.get("PP_SLAB_PARANOID"))  # check the halo populations after every rebuild
HALO_TRANSPORT = os.environ.get("PP_SLAB_HALO", "peer")  # "peer": stores into the neighbour's memory over NVLink; "nccl": send/recv


def _dbg(rank, *a):
    if _DEBUG:
        print(f"[slab {rank}]", *a, file=sys.stderr, flush=True)


def layer_partition(n_layers, world):
    """[(x0, x1)] per rank: whole pairs of layers, the first ranks take the remainder (as pp_upload does)."""
    pairs = n_layers // 2
    base, rem = divmod(pairs, world)
    if base < 1:
        raise ValueError(f"{n_layers} x layers cannot be split over {world} engines (at least 2 layers each)")
    out = []
    for r in range(world):
        x0 = 2 * (r * base + min(r, rem))
        out.append((x0, x0 + 2 * (base + (1 if r < rem else 0))))
    return out


def neighbours(rank, world):
    return (rank - 1) % world, (rank + 1) % world


class _DeviceBuffer:
    """Zero-copy torch view of library-owned device memory (CUDA array interface)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def device_view(ptr, n, dtype, device):
    if n == 0:
        return torch.empty(0, dtype=dtype, device=device)
    if device.type == "cpu":  # stand-in engines of the CPU tests hand out host pointers
        import ctypes
        itemsize = {torch.float64: 8, torch.float32: 4, torch.int32: 4}[dtype]
        return torch.frombuffer((ctypes.c_char * (n * itemsize)).from_address(ptr), dtype=dtype)
    typestr = {torch.float64: "<f8", torch.float32: "<f4", torch.int32: "<i4"}[dtype]
    return torch.as_tensor(_DeviceBuffer(ptr, n, typestr), device=device)


def exchange_with_neighbours(rank, world, to_left, to_right, from_left, from_right, group=None):
    """Sends ``to_left`` / ``to_right`` to the ring neighbours and fills ``from_left`` / ``from_right``.
    Lists of tensors are sent in order.  With two ranks both neighbours are the same peer: receives are
    posted right-first so that a peer's to-left message lands in this rank's from-right buffer."""
    left, right = neighbours(rank, world)
    ops = []
    for t in to_left:
        if t.numel():
            ops.append(dist.P2POp(dist.isend, t, left, group))
    for t in to_right:
        if t.numel():
            ops.append(dist.P2POp(dist.isend, t, right, group))
    for t in from_right:
        if t.numel():
            ops.append(dist.P2POp(dist.irecv, t, right, group))
    for t in from_left:
        if t.numel():
            ops.append(dist.P2POp(dist.irecv, t, left, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def exchange_counts(rank, world, n_to_left, n_to_right, device, group=None):
    """-> (n_from_left, n_from_right)"""
    send_l = torch.tensor([n_to_left], dtype=torch.int64, device=device)
    send_r = torch.tensor([n_to_right], dtype=torch.int64, device=device)
    recv_l = torch.zeros(1, dtype=torch.int64, device=device)
    recv_r = torch.zeros(1, dtype=torch.int64, device=device)
    exchange_with_neighbours(rank, world, [send_l], [send_r], [recv_l], [recv_r], group)
    return int(recv_l.item()), int(recv_r.item())


class SlabEngine:
    """One rank of a slab-decomposed simulation.  Mirrors :class:`Engine` for the NVT path."""

    def __init__(self, rank, world, device_index, engine=None, transport=None, **params):
        """``engine`` / ``transport`` are for the CPU tests of this driver: a stand-in engine object (tensors then live
        on the CPU) and "peer" / "nccl" instead of the PP_SLAB_HALO environment variable."""
        if world < 2:
            raise ValueError("SlabEngine needs at least two ranks; use Engine on one GPU")
        self.rank, self.world = rank, world
        if engine is None:
            self.device = torch.device("cuda", device_index)
            self.eng = Engine(device=device_index, **params)
            self.eng.set_stream(torch.cuda.current_stream(self.device).cuda_stream)
        else:
            self.device = torch.device("cpu")
            self.eng = engine
        self.eng.slab_setup(rank, world)
        self.N = 0
        self._halo_cache = [None, None]
        self._rebuilds = 0
        self.peer_halo = (transport or HALO_TRANSPORT) == "peer"

    def close(self):
        if self.peer_halo:
            self.eng.sync()
            dist.barrier()  # nobody unmaps while a neighbour may still store into it
            self.eng.slab_peer_close()
            dist.barrier()
        self.eng.close()

    def box_move(self, kind, u):
        """One MC_change_volume (kind 6) / MC_change_Lx (kind 7) under the decomposition: every engine rescales and counts
        its own layers, {bonds, overlap} are summed over the ranks (16 bytes), every engine takes the same decision.
        -> 1 accepted, 0 rejected, -1 rejected by the Lyz window."""
        running = self._allreduce([int(round(-self.eng.observables().energy))])[0]
        b, ov, status = self.eng.slab_box_trial(kind, u, running)
        if status < 0:
            return -1 if status == -1 else 0
        if status == 2:
            raise NotImplementedError("a trial box that needs another cell grid re-partitions the slabs: drive this run through pp_group "
                                      "(host/patchy_b200 with GPUs = n), which gathers and redistributes")
        b, ov = self._allreduce([b, ov])
        return self.eng.slab_box_finish(b, int(ov > 0))

    def upload(self, box, r, orient):
        """Every rank is handed the whole configuration and takes only the particles of the layers it holds (owned + halo):
        pp_upload_select, on the GPU when the arrays are pinned host memory."""
        import time
        timing = os.environ.get("PP_SLAB_TIMING") and self.rank == 0
        t = [time.perf_counter()]
        r = np.asarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.asarray(orient, dtype=np.float64).reshape(-1, 9)
        self.N = int(r.shape[0])
        if self.peer_halo:
            self.eng.sync()
            if getattr(self, "_connected", False):
                dist.barrier()  # nobody unmaps or frees while a neighbour may still store into it
                self.eng.slab_peer_close()
                dist.barrier()
                self._connected = False
        t.append(time.perf_counter())
        if hasattr(self.eng, "upload_select"):
            self.eng.upload_select(box, r, orient)
            self.h2d_bytes = int(r.nbytes + 108 * self.eng.slab_range()[3])  # every position + the selected particles' state
        else:  # stand-in engines of the CPU tests
            self.eng.upload(box, r, orient)
            self.h2d_bytes = 0
        t.append(time.perf_counter())
        self._halo_cache = [None, None]
        self._rebuilds = 0
        if self.peer_halo:
            self._connect_peers()
        t.append(time.perf_counter())
        if timing:
            print(f"[slab upload] close {1e3 * (t[1] - t[0]):.2f} ms, select + copy + cell build {1e3 * (t[2] - t[1]):.2f} ms, "
                  f"connect {1e3 * (t[3] - t[2]):.2f} ms", file=sys.stderr, flush=True)

    def _connect_peers(self):
        """Every engine exports the description of its device state (CUDA IPC) and maps those of its ring neighbours."""
        mine = self.eng.slab_peer_export()
        infos = [None] * self.world
        dist.all_gather_object(infos, mine)
        left, right = neighbours(self.rank, self.world)
        self.eng.slab_peer_connect(infos[left], infos[right])
        self._connected = True
        dist.barrier()  # nobody stores into a neighbour before everybody has mapped

    def _allreduce(self, values, op=dist.ReduceOp.SUM):
        t = torch.tensor(values, dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=op)
        return [int(x) for x in t.tolist()]

    def total_bonds(self):
        b, ov = self.eng.total_bonds()
        b, ov = self._allreduce([b, ov])
        return b, int(ov > 0)

    def observables(self, reset=False):
        o = self.eng.observables(reset=reset)
        bonds, tries, acc, conf, ovl, checks = self._allreduce(
            [int(round(-o.energy)), o.tries[0], o.accepted[0], o.confined, o.overlapped, o.pair_checks])
        return {"energy": -float(bonds), "tries": tries, "accepted": acc, "confined": conf, "overlapped": ovl,
                "pair_checks": checks, "step": o.step, "N": self.N}

    def check_energy(self):
        """MC_check_energy across the slabs: sum of running energies == sum of recounts, no overlap."""
        running = int(round(-self.eng.observables().energy))
        b, ov, ok_local = self.eng.check_energy()
        b, ov, running, bad = self._allreduce([b, ov, running, 0 if ok_local else 1])
        return b, int(ov > 0), (b == running and ov == 0 and bad == 0)

    def _halo_views(self, send_side):
        """Zero-copy views of the layer slices of one halo refresh; they only change when the cells are rebuilt."""
        e, dev = self.eng, self.device
        sp, so, sf, ns = e.slab_halo(send_side, recv=False)
        rp, ro, rf, nr = e.slab_halo(1 - send_side, recv=True)
        send = [device_view(sp, 4 * ns, torch.float64, dev), device_view(so, 9 * ns, torch.float64, dev),
                device_view(sf, 4 * ns, torch.int32, dev)]
        recv = [device_view(rp, 4 * nr, torch.float64, dev), device_view(ro, 9 * nr, torch.float64, dev),
                device_view(rf, 4 * nr, torch.int32, dev)]
        return send, recv

    def _halo_refresh(self, send_side):
        """send / recv transport only.  send_side 0: first owned layer -> left neighbour (its right halo); 1: last owned layer -> right."""
        if self._halo_cache[send_side] is None:
            self._halo_cache[send_side] = self._halo_views(send_side)
        send, recv = self._halo_cache[send_side]
        if send_side == 0:
            exchange_with_neighbours(self.rank, self.world, send, [], [], recv)
        else:
            exchange_with_neighbours(self.rank, self.world, [], send, recv, [])

    def _check_halo_counts(self):
        """Once per shift period: both sides of every slab boundary must hold the same layer populations
        (a mismatch would otherwise deadlock the fixed-size halo refresh)."""
        e = self.eng
        first, last = e.slab_halo(0, recv=False)[3], e.slab_halo(1, recv=False)[3]
        halo_l, halo_r = e.slab_halo(0, recv=True)[3], e.slab_halo(1, recv=True)[3]
        from_l, from_r = exchange_counts(self.rank, self.world, first, last, self.device)
        # my first layer went to the left neighbour, so what I got from the right is ITS first layer = my right halo
        if from_r != halo_r or from_l != halo_l:
            raise RuntimeError(f"rank {self.rank}: halo populations differ across a slab boundary: right halo {halo_r} vs owner {from_r}, "
                               f"left halo {halo_l} vs owner {from_l}")

    def _migrate(self):
        """Start of a shift period: particles change owner, halos are rebuilt from records (send/recv)."""
        e, dev = self.eng, self.device
        (pl, nl), (pr, nr) = e.slab_outbox(0), e.slab_outbox(1)
        _dbg(self.rank, "exchange: to left", nl, "to right", nr)
        n_from_l, n_from_r = exchange_counts(self.rank, self.world, nl, nr, dev)
        _dbg(self.rank, "exchange: from left", n_from_l, "from right", n_from_r)
        il, ir = e.slab_inbox(0, n_from_l), e.slab_inbox(1, n_from_r)
        exchange_with_neighbours(
            self.rank, self.world,
            [device_view(pl, RECORD_DOUBLES * nl, torch.float64, dev)], [device_view(pr, RECORD_DOUBLES * nr, torch.float64, dev)],
            [device_view(il, RECORD_DOUBLES * n_from_l, torch.float64, dev)], [device_view(ir, RECORD_DOUBLES * n_from_r, torch.float64, dev)])
        e.slab_sweep_rebuild()
        self._after_rebuild()

    def _after_rebuild(self):
        _dbg(self.rank, "rebuilt", self.eng.slab_range())
        self._halo_cache = [None, None]
        self._rebuilds += 1
        if self._rebuilds <= 4 or self._rebuilds % 16 == 0 or PARANOID:  # the check costs a host synchronisation
            self._check_halo_counts()

    def sweep(self, nsteps=1, sync=True):
        e = self.eng
        left = int(nsteps)
        while left > 0:
            if self.peer_halo:
                # the data path is inside the library: accepted moves of the boundary layers are stored straight into the
                # neighbour's halo by the sweep kernels, neighbours exchange phase numbers on the device; the driver only
                # starts each shift period (migration) and lets the library run on until the next one
                if e.slab_migrate():
                    self._after_rebuild()
                done = e.slab_sweeps(left)
            else:
                if e.slab_sweep_begin():
                    self._migrate()
                done = 1
                e.slab_pass_block(0)
                self._halo_refresh(0)
                e.slab_pass_block(1)
                self._halo_refresh(1)
                e.slab_sweep_end()
            left -= done
        if sync:
            e.sync()

    def download(self):
        """-> (r, orient) of the whole system on every rank, in the caller's particle order."""
        ids, r, m = self.eng.slab_download()
        counts = [None] * self.world
        dist.all_gather_object(counts, len(ids))
        cap = max(counts)
        pad = lambda a, w: np.concatenate([a.reshape(len(ids), w), np.zeros((cap - len(ids), w))]) if cap > len(ids) else a.reshape(len(ids), w)
        packed = torch.from_numpy(np.concatenate([pad(ids.astype(np.float64), 1), pad(r, 3), pad(m, 9)], axis=1)).to(self.device)
        out = [torch.empty_like(packed) for _ in range(self.world)]
        dist.all_gather(out, packed)
        R = np.zeros((self.N, 3))
        M = np.zeros((self.N, 3, 3))
        seen = np.zeros(self.N, dtype=bool)
        for k, t in enumerate(out):
            a = t.cpu().numpy()[:counts[k]]
            i = a[:, 0].astype(np.int64)
            R[i] = a[:, 1:4]
            M[i] = a[:, 4:13].reshape(-1, 3, 3)
            seen[i] = True
        if not seen.all():
            raise RuntimeError(f"{(~seen).sum()} particles are owned by no slab")
        return R, M
