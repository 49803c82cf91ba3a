"""ctypes mirror of include/pp_b200.h (the drop-in C ABI).

Names follow the reference's vocabulary: an :class:`Engine` plays the role of ``System`` plus its
``do_ensemble`` plugin (defs.h:58-112, MC.c:85-131); ``sweep`` is one or more MC steps
(main.c:119); ``total_bonds`` is the from-scratch energy of main.c:55-61.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
N_MOVES = 8
ROTO_TRASL, VOLUME, LX = 0, 6, 7
NVT, NPT = 0, 5
K_CELLS, K_SWEEP, K_BONDS, K_RESCALE, K_PROPOSALS, K_ORTHO, K_RNG = range(7)
KERNEL_CLASSES = ["cells", "sweep", "bonds", "rescale", "proposals", "ortho", "rng", "halo"]


class LibraryMissing(RuntimeError):
    pass


class Params(C.Structure):
    """pp_params: the input-file keys of the hot path (system.c:61-196)."""
    _fields_ = [("ensemble", C.c_int), ("temperature", C.c_double), ("disp_max", C.c_double),
                ("theta_max", C.c_double), ("kf_delta", C.c_double), ("kf_cosmax", C.c_double),
                ("P", C.c_double), ("rescale_factor_max", C.c_double), ("Lx_move", C.c_int),
                ("Lx_change_max", C.c_double), ("Lyz_min", C.c_double), ("Lyz_max", C.c_double),
                ("Lx_reference_sign", C.c_int), ("seed", C.c_uint64), ("device", C.c_int),
                ("lanes_per_cell", C.c_int), ("cell_lanes", C.c_int), ("cells_max_occupancy", C.c_int),
                ("shift_every", C.c_int), ("cells_mid_occupancy", C.c_int), ("x_layers_multiple", C.c_int)]


class Observables(C.Structure):
    """pp_observables: what output_print reads from System (output.c:119-196)."""
    _fields_ = [("step", C.c_longlong), ("N", C.c_int), ("box", C.c_double * 3), ("energy", C.c_double),
                ("tries", C.c_longlong * N_MOVES), ("accepted", C.c_longlong * N_MOVES),
                ("confined", C.c_longlong), ("overlapped", C.c_longlong), ("acceptance_roto", C.c_double),
                ("pair_checks", C.c_longlong), ("n_side", C.c_int * 3),
                ("volume", C.c_double), ("w_tries", C.c_ulonglong), ("w_accepted", C.c_ulonglong)]


def library_path():
    return os.path.join(_HERE, "libpatchy_b200.so")


_lib = None
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")

# every symbol include/pp_b200.h declares
ABI_SYMBOLS = ["pp_default_params", "pp_abi_version", "pp_create", "pp_destroy", "pp_last_error", "pp_upload",
               "pp_download", "pp_total_bonds", "pp_bond_counts", "pp_bond_lists", "pp_sweep", "pp_sync", "pp_get_observables",
               "pp_check_energy", "pp_eval_proposals", "pp_volume_move", "pp_Lx_move", "pp_set_stream",
               "pp_set_profiling", "pp_kernel_stats", "pp_reset_kernel_stats", "pp_slab_setup", "pp_slab_range",
               "pp_slab_sweep_begin", "pp_slab_outbox", "pp_slab_inbox", "pp_slab_sweep_rebuild", "pp_slab_pass_block",
               "pp_slab_halo", "pp_slab_sweep_end", "pp_slab_download", "pp_slab_peer_export", "pp_slab_peer_connect",
               "pp_slab_peer_close", "pp_slab_sweeps", "pp_slab_migrate", "pp_upload_subset", "pp_upload_select", "pp_slab_select",
               "pp_slab_box_trial", "pp_slab_box_finish", "pp_group_create", "pp_group_destroy", "pp_group_last_error",
               "pp_group_upload", "pp_group_download", "pp_group_total_bonds", "pp_group_bond_lists", "pp_group_sweep",
               "pp_group_sync", "pp_group_get_observables", "pp_group_check_energy", "pp_group_volume_move",
               "pp_group_Lx_move", "pp_group_engine", "pp_reserve", "pp_particle_bonds", "pp_probe_bonds", "pp_add_remove", "pp_gc_step", "pp_patch_neighbours",
               "pp_range_neighbours"]


class GCResult(C.Structure):
    """pp_gc_result (include/pp_b200.h)"""
    _fields_ = [("kind", C.c_int), ("accepted", C.c_int), ("overlap", C.c_int), ("bonds", C.c_int), ("index", C.c_int),
                ("n_uniforms", C.c_int), ("N", C.c_int)]


class PeerInfo(C.Structure):
    """pp_peer_info: one engine's device state as a neighbour in another process maps it (CUDA IPC handle + offsets)."""
    _fields_ = [("ipc", C.c_ubyte * 64), ("off_pos", C.c_ulonglong * 3), ("off_ori", C.c_ulonglong * 2),
                ("off_posf", C.c_ulonglong), ("off_mig", C.c_ulonglong * 2), ("off_ctrl", C.c_ulonglong),
                ("cap_records", C.c_longlong)]


def load_library():
    """Loads libpatchy_b200.so; raises LibraryMissing if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise LibraryMissing(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "or `make -C patchyparticles_b200/csrc`. There is no CPU fallback.")
    L = C.CDLL(path)
    H = C.c_void_p
    L.pp_default_params.argtypes = [C.POINTER(Params)]
    L.pp_default_params.restype = None
    L.pp_create.argtypes = [C.POINTER(Params), C.POINTER(H)]
    L.pp_destroy.argtypes = [H]
    L.pp_last_error.argtypes = [H]
    L.pp_last_error.restype = C.c_char_p
    L.pp_upload.argtypes = [H, C.c_int, _dp, _dp, _dp]
    L.pp_download.argtypes = [H, _dp, _dp]
    L.pp_total_bonds.argtypes = [H, C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    L.pp_bond_counts.argtypes = [H, _ip]
    L.pp_bond_lists.argtypes = [H, C.c_int, _ip, _ip]
    L.pp_sweep.argtypes = [H, C.c_longlong]
    L.pp_sync.argtypes = [H]
    L.pp_get_observables.argtypes = [H, C.POINTER(Observables), C.c_int]
    L.pp_check_energy.argtypes = [H, C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.pp_eval_proposals.argtypes = [H, C.c_int, _ip, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip]
    L.pp_volume_move.argtypes = [H, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_int)]
    L.pp_Lx_move.argtypes = [H, C.c_double, C.c_double, C.POINTER(C.c_int)]
    L.pp_set_stream.argtypes = [H, C.c_void_p]
    L.pp_set_profiling.argtypes = [H, C.c_int]
    L.pp_kernel_stats.argtypes = [H, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_double)]
    L.pp_reset_kernel_stats.argtypes = [H]
    PI, PV = C.POINTER(C.c_int), C.POINTER(C.c_void_p)
    L.pp_slab_setup.argtypes = [H, C.c_int, C.c_int]
    L.pp_slab_range.argtypes = [H, PI, PI, PI, PI]
    L.pp_slab_sweep_begin.argtypes = [H, PI]
    L.pp_slab_outbox.argtypes = [H, C.c_int, PV, PI]
    L.pp_slab_inbox.argtypes = [H, C.c_int, C.c_int, PV]
    L.pp_slab_sweep_rebuild.argtypes = [H]
    L.pp_slab_pass_block.argtypes = [H, C.c_int]
    L.pp_slab_halo.argtypes = [H, C.c_int, C.c_int, PV, PV, PV, PI]
    L.pp_slab_sweep_end.argtypes = [H]
    L.pp_slab_peer_export.argtypes = [H, C.POINTER(PeerInfo)]
    L.pp_slab_peer_connect.argtypes = [H, C.POINTER(PeerInfo), C.POINTER(PeerInfo)]
    L.pp_slab_peer_close.argtypes = [H]
    L.pp_upload_subset.argtypes = [H, C.c_int, C.c_int, _ip, _dp, _dp, _dp]
    L.pp_upload_select.argtypes = [H, C.c_int, _dp, _dp, _dp]
    L.pp_slab_select.argtypes = [C.POINTER(Params), C.c_int, C.c_int, C.c_int, _dp, _dp,
                                 np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")]
    L.pp_slab_box_trial.argtypes = [H, C.c_int, _dp, C.c_longlong, C.POINTER(C.c_longlong), PI, PI]
    L.pp_slab_box_finish.argtypes = [H, C.c_longlong, C.c_int, PI]
    G = C.c_void_p
    L.pp_group_create.argtypes = [C.POINTER(Params), C.c_int, _ip, C.POINTER(G)]
    L.pp_group_destroy.argtypes = [G]
    L.pp_group_last_error.argtypes = [G]
    L.pp_group_last_error.restype = C.c_char_p
    L.pp_group_upload.argtypes = [G, C.c_int, _dp, _dp, _dp]
    L.pp_group_download.argtypes = [G, _dp, _dp]
    L.pp_group_total_bonds.argtypes = [G, C.POINTER(C.c_longlong), PI]
    L.pp_group_bond_lists.argtypes = [G, C.c_int, _ip, _ip]
    L.pp_group_sweep.argtypes = [G, C.c_longlong]
    L.pp_group_sync.argtypes = [G]
    L.pp_group_get_observables.argtypes = [G, C.POINTER(Observables), C.c_int]
    L.pp_group_check_energy.argtypes = [G, C.POINTER(C.c_longlong), PI, PI]
    L.pp_group_volume_move.argtypes = [G, C.c_double, C.c_double, C.c_double, PI]
    L.pp_group_Lx_move.argtypes = [G, C.c_double, C.c_double, PI]
    L.pp_group_engine.argtypes = [G, C.c_int, C.POINTER(H)]
    L.pp_slab_sweeps.argtypes = [H, C.c_int, PI]
    L.pp_slab_migrate.argtypes = [H, PI]
    L.pp_slab_download.argtypes = [H, PI, _ip, _dp, _dp]
    L.pp_reserve.argtypes = [H, C.c_int]
    L.pp_particle_bonds.argtypes = [H, C.c_int, PI, PI]
    L.pp_probe_bonds.argtypes = [H, _dp, _dp, PI, PI]
    L.pp_add_remove.argtypes = [H, C.c_double, C.c_int, C.c_int, _dp, C.c_int, C.POINTER(GCResult)]
    L.pp_gc_step.argtypes = [H, C.c_double, C.c_int, C.c_int, C.c_void_p, PI, PI]
    L.pp_patch_neighbours.argtypes = [H, C.c_int, _ip, _ip, _ip]
    L.pp_range_neighbours.argtypes = [H, C.c_int, _ip, C.c_int, _ip, _ip]
    _lib = L
    return L


class EngineError(RuntimeError):
    pass


class Engine:
    """One GPU's Monte Carlo engine (pp_handle)."""

    def __init__(self, ensemble=NVT, temperature=0.1, disp_max=0.1, theta_max=0.1, kf_delta=0.119, kf_cosmax=0.98,
                 P=0.0, rescale_factor_max=0.0, Lx_move=0, Lx_change_max=0.0, Lyz_min=0.0, Lyz_max=0.0,
                 Lx_reference_sign=1, seed=12345, device=0, lanes_per_cell=0, cell_lanes=0, shift_every=0,
                 cells_max_occupancy=-1, cells_mid_occupancy=-1, x_layers_multiple=0):
        self.L = load_library()
        p = Params()
        self.L.pp_default_params(C.byref(p))
        p.ensemble, p.temperature, p.disp_max, p.theta_max = ensemble, temperature, disp_max, theta_max
        p.kf_delta, p.kf_cosmax, p.P, p.rescale_factor_max = kf_delta, kf_cosmax, P, rescale_factor_max
        p.Lx_move, p.Lx_change_max, p.Lyz_min, p.Lyz_max = Lx_move, Lx_change_max, Lyz_min, Lyz_max
        p.Lx_reference_sign, p.seed, p.device = Lx_reference_sign, seed, device
        p.lanes_per_cell, p.cell_lanes, p.shift_every = lanes_per_cell, cell_lanes, shift_every
        p.cells_max_occupancy = cells_max_occupancy
        p.cells_mid_occupancy = cells_mid_occupancy
        p.x_layers_multiple = x_layers_multiple
        self.params = p
        self.h = C.c_void_p()
        rc = self.L.pp_create(C.byref(p), C.byref(self.h))
        if rc != 0:
            msg = self.L.pp_last_error(self.h).decode() if self.h else "pp_create failed"
            if self.h:
                self.L.pp_destroy(self.h)
            self.h = None
            raise EngineError(msg)
        self.N = 0

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.L.pp_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.pp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- state ----------------------------------------------------------------------------
    def upload(self, box, r, orient):
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.ascontiguousarray(orient, dtype=np.float64).reshape(-1, 9)
        box = np.ascontiguousarray(box, dtype=np.float64)
        self.N = int(r.shape[0])
        if self.N == 0:
            r, orient = np.zeros((1, 3)), np.zeros((1, 9))
        self._check(self.L.pp_upload(self.h, self.N, box, r, orient))

    def download(self, out=None):
        """positions and orientation matrices in the caller's particle order; out = (r[N, 3], orient[N, 9]) float64 arrays to fill
        (pinned host memory makes the copy a DMA at PCIe speed; fresh pageable arrays otherwise)"""
        if out is not None:
            r, m = out
            if r.shape[0] < self.N or m.shape[0] < self.N or r.dtype != np.float64 or m.dtype != np.float64:
                raise ValueError("download: output arrays too small or not float64")
        else:
            r = np.zeros((max(self.N, 1), 3))
            m = np.zeros((max(self.N, 1), 9))
        self._check(self.L.pp_download(self.h, r, m))
        return r[:self.N], m[:self.N].reshape(-1, 3, 3)

    # ---- energies --------------------------------------------------------------------------
    def total_bonds(self):
        b, ov = C.c_longlong(0), C.c_int(0)
        self._check(self.L.pp_total_bonds(self.h, C.byref(b), C.byref(ov)))
        return int(b.value), int(ov.value)

    def bond_counts(self):
        c = np.zeros(max(self.N, 1), dtype=np.int32)
        self._check(self.L.pp_bond_counts(self.h, c))
        return c[:self.N]

    def bond_lists(self, max_per_particle=8):
        """-> (counts[N], neighbours[N, max_per_particle]) as output_print_bonds lists them (output.c:213-268); -1 = unused"""
        c = np.zeros(max(self.N, 1), dtype=np.int32)
        nb = np.full((max(self.N, 1), int(max_per_particle)), -1, dtype=np.int32)
        self._check(self.L.pp_bond_lists(self.h, int(max_per_particle), c, nb.reshape(-1)))
        return c[:self.N], nb[:self.N]

    def check_energy(self):
        b, ov, ok = C.c_longlong(0), C.c_int(0), C.c_int(0)
        self._check(self.L.pp_check_energy(self.h, C.byref(b), C.byref(ov), C.byref(ok)))
        return int(b.value), int(ov.value), bool(ok.value)

    # ---- moves -----------------------------------------------------------------------------
    def sweep(self, nsteps=1, sync=True):
        self._check(self.L.pp_sweep(self.h, int(nsteps)))
        if sync:
            self._check(self.L.pp_sync(self.h))

    def sync(self):
        self._check(self.L.pp_sync(self.h))

    def observables(self, reset=False):
        o = Observables()
        self._check(self.L.pp_get_observables(self.h, C.byref(o), int(reset)))
        return o

    def eval_proposals(self, idx, disp, orient_new, u, commit=False):
        n = len(idx)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        disp = np.ascontiguousarray(disp, dtype=np.float64).reshape(-1, 3)
        orient_new = np.ascontiguousarray(orient_new, dtype=np.float64).reshape(-1, 9)
        u = np.ascontiguousarray(u, dtype=np.float64)
        dE = np.zeros(max(n, 1), dtype=np.int32)
        ov = np.zeros(max(n, 1), dtype=np.int32)
        acc = np.zeros(max(n, 1), dtype=np.int32)
        if n:
            self._check(self.L.pp_eval_proposals(self.h, n, idx, disp, orient_new, u, int(commit), dE, ov, acc))
        return dE[:n], ov[:n], acc[:n]

    def volume_move(self, u_dir, u_vol, u_acc):
        a = C.c_int(0)
        self._check(self.L.pp_volume_move(self.h, u_dir, u_vol, u_acc, C.byref(a)))
        return int(a.value)

    def Lx_move(self, u_lx, u_acc):
        a = C.c_int(0)
        self._check(self.L.pp_Lx_move(self.h, u_lx, u_acc, C.byref(a)))
        return int(a.value)

    # ---- grand-canonical insertion / removal (MC_add_remove, MC.c:342-448) ------------------------
    def reserve(self, n_max):
        """room for up to n_max particles; before upload"""
        self._check(self.L.pp_reserve(self.h, int(n_max)))

    def particle_bonds(self, idx):
        b, ov = C.c_int(0), C.c_int(0)
        self._check(self.L.pp_particle_bonds(self.h, int(idx), C.byref(b), C.byref(ov)))
        return int(b.value), int(ov.value)

    def probe_bonds(self, r, orient):
        b, ov = C.c_int(0), C.c_int(0)
        self._check(self.L.pp_probe_bonds(self.h, np.ascontiguousarray(r, dtype=np.float64).reshape(3),
                                          np.ascontiguousarray(orient, dtype=np.float64).reshape(9), C.byref(b), C.byref(ov)))
        return int(b.value), int(ov.value)

    def add_remove(self, activity, n_min, n_max, uniforms):
        """one MC_add_remove with recorded uniforms -> GCResult"""
        u = np.ascontiguousarray(uniforms, dtype=np.float64).reshape(-1)
        res = GCResult()
        self._check(self.L.pp_add_remove(self.h, float(activity), int(n_min), int(n_max), u, int(u.size), C.byref(res)))
        self.N = int(res.N)
        return res

    def gc_step(self, activity, n_min, n_max, hist=None):
        """the insertion / removal trials of one MC step -> (trials, full); hist: int64 array of n_max - n_min + 1 bins (SUS)"""
        n, full = C.c_int(0), C.c_int(0)
        hp = hist.ctypes.data_as(C.c_void_p) if hist is not None else None
        self._check(self.L.pp_gc_step(self.h, float(activity), int(n_min), int(n_max), hp, C.byref(n), C.byref(full)))
        o = Observables()
        self._check(self.L.pp_get_observables(self.h, C.byref(o), 0))
        self.N = int(o.N)
        return int(n.value), int(full.value)

    # ---- neighbour queries of the cluster moves (avb.c:43-75, vmmc.c:148-229) ----------------------
    def patch_neighbours(self, idx):
        """-> (neighbours[n, 4]: bonded neighbour per patch or -1, num_neighbours[n])"""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        nb = np.full((max(len(idx), 1), 4), -1, dtype=np.int32)
        num = np.zeros(max(len(idx), 1), dtype=np.int32)
        self._check(self.L.pp_patch_neighbours(self.h, len(idx), idx, nb.reshape(-1), num))
        return nb[:len(idx)], num[:len(idx)]

    def range_neighbours(self, idx, max_per_particle=32):
        """-> (counts[n], neighbours[n, max_per_particle]: particles with dist2 < (1 + delta)^2, -1 beyond the count)"""
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        nb = np.full((max(len(idx), 1), max_per_particle), -1, dtype=np.int32)
        cnt = np.zeros(max(len(idx), 1), dtype=np.int32)
        self._check(self.L.pp_range_neighbours(self.h, len(idx), idx, int(max_per_particle), cnt, nb.reshape(-1)))
        return cnt[:len(idx)], nb[:len(idx)]

    # ---- x-slab decomposition hooks (see slab.py) ------------------------------------------------
    def slab_setup(self, rank, world):
        self._check(self.L.pp_slab_setup(self.h, rank, world))

    def slab_range(self):
        v = [C.c_int(0) for _ in range(4)]
        self._check(self.L.pp_slab_range(self.h, *[C.byref(x) for x in v]))
        return tuple(int(x.value) for x in v)  # x0, x1, nx, n_local

    def slab_sweep_begin(self):
        ex = C.c_int(0)
        self._check(self.L.pp_slab_sweep_begin(self.h, C.byref(ex)))
        return bool(ex.value)

    def slab_outbox(self, side):
        ptr, n = C.c_void_p(), C.c_int(0)
        self._check(self.L.pp_slab_outbox(self.h, side, C.byref(ptr), C.byref(n)))
        return ptr.value, int(n.value)

    def slab_inbox(self, side, count):
        ptr = C.c_void_p()
        self._check(self.L.pp_slab_inbox(self.h, side, int(count), C.byref(ptr)))
        return ptr.value

    def slab_sweep_rebuild(self):
        self._check(self.L.pp_slab_sweep_rebuild(self.h))

    def slab_pass_block(self, cx):
        self._check(self.L.pp_slab_pass_block(self.h, cx))

    def slab_halo(self, side, recv):
        pos, ori, posf, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int(0)
        self._check(self.L.pp_slab_halo(self.h, side, int(recv), C.byref(pos), C.byref(ori), C.byref(posf), C.byref(n)))
        return pos.value, ori.value, posf.value, int(n.value)

    def slab_sweep_end(self):
        self._check(self.L.pp_slab_sweep_end(self.h))

    def slab_peer_export(self):
        """-> pp_peer_info of this engine's device state (CUDA IPC handle + offsets) as bytes"""
        info = PeerInfo()
        self._check(self.L.pp_slab_peer_export(self.h, C.byref(info)))
        return bytes(info)

    def slab_peer_connect(self, left_info, right_info):
        l, r = PeerInfo.from_buffer_copy(left_info), PeerInfo.from_buffer_copy(right_info)
        self._check(self.L.pp_slab_peer_connect(self.h, C.byref(l), C.byref(r)))

    def upload_subset(self, N_global, ids, box, r, orient):
        """pp_upload_subset: the particles of the layers this slab engine holds, with their global ids"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.ascontiguousarray(orient, dtype=np.float64).reshape(-1, 9)
        self.N = int(N_global)
        n = len(ids)
        if n == 0:
            ids, r, orient = np.zeros(1, dtype=np.int32), np.zeros((1, 3)), np.zeros((1, 9))
        self._check(self.L.pp_upload_subset(self.h, int(N_global), n, ids, np.ascontiguousarray(box, dtype=np.float64), r, orient))

    def upload_select(self, box, r, orient):
        """pp_upload_select: the whole configuration; a slab engine takes the particles of its own layers (on the GPU when
        the arrays are pinned host memory)"""
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.ascontiguousarray(orient, dtype=np.float64).reshape(-1, 9)
        self.N = int(r.shape[0])
        if self.N == 0:
            r, orient = np.zeros((1, 3)), np.zeros((1, 9))
        self._check(self.L.pp_upload_select(self.h, self.N, np.ascontiguousarray(box, dtype=np.float64), r, orient))

    def slab_select(self, rank, world, box, r):
        """-> boolean mask of the particles slab `rank` of `world` has to receive (pp_slab_select)"""
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        mine = np.zeros(max(len(r), 1), dtype=np.uint8)
        if self.L.pp_slab_select(C.byref(self.params), rank, world, len(r), np.ascontiguousarray(box, dtype=np.float64), r, mine) != 0:
            raise EngineError(f"the box cannot be split into {world} slabs of at least two cell layers")
        return mine[:len(r)].astype(bool)

    def slab_box_trial(self, kind, u, bonds_running_total):
        """-> (bonds_local, overlap_local, status); see pp_slab_box_trial"""
        b, ov, st = C.c_longlong(0), C.c_int(0), C.c_int(0)
        uu = np.zeros(3)
        uu[:len(u)] = u
        self._check(self.L.pp_slab_box_trial(self.h, int(kind), uu, int(bonds_running_total), C.byref(b), C.byref(ov), C.byref(st)))
        return int(b.value), int(ov.value), int(st.value)

    def slab_box_finish(self, bonds_total, overlap_any):
        a = C.c_int(0)
        self._check(self.L.pp_slab_box_finish(self.h, int(bonds_total), int(overlap_any), C.byref(a)))
        return int(a.value)

    def slab_sweeps(self, max_sweeps):
        """-> sweeps done before the next grid shift (peer connection only)"""
        done = C.c_int(0)
        self._check(self.L.pp_slab_sweeps(self.h, int(max_sweeps), C.byref(done)))
        return int(done.value)

    def slab_migrate(self):
        """Start of a shift period over peer memory -> True if the cells were rebuilt."""
        done = C.c_int(0)
        self._check(self.L.pp_slab_migrate(self.h, C.byref(done)))
        return bool(done.value)

    def slab_peer_close(self):
        self._check(self.L.pp_slab_peer_close(self.h))

    def slab_download(self):
        n = C.c_int(0)
        cap = max(self.N, 1)
        ids = np.zeros(cap, dtype=np.int32)
        r = np.zeros((cap, 3))
        m = np.zeros((cap, 9))
        self._check(self.L.pp_slab_download(self.h, C.byref(n), ids, r, m))
        k = int(n.value)
        return ids[:k], r[:k], m[:k].reshape(-1, 3, 3)

    # ---- measurement plumbing --------------------------------------------------------------
    def set_stream(self, cuda_stream):
        self._check(self.L.pp_set_stream(self.h, C.c_void_p(cuda_stream)))

    def set_profiling(self, enable):
        self._check(self.L.pp_set_profiling(self.h, int(bool(enable))))

    def kernel_stats(self):
        out = {}
        for k, name in enumerate(KERNEL_CLASSES):
            n, ms = C.c_longlong(0), C.c_double(0)
            self._check(self.L.pp_kernel_stats(self.h, k, C.byref(n), C.byref(ms)))
            out[name] = (int(n.value), float(ms.value))
        return out

    def reset_kernel_stats(self):
        self._check(self.L.pp_reset_kernel_stats(self.h))


class Group:
    """pp_group: one simulation in n x-slabs driven from this thread; devices may repeat (several slabs on one GPU).
    Same calls as :class:`Engine`; follows the single engine's trajectory bit for bit."""

    def __init__(self, n, devices=None, **params):
        self.L = load_library()
        p = Params()
        self.L.pp_default_params(C.byref(p))
        for k, v in params.items():
            if not hasattr(p, k):
                raise TypeError(f"unknown engine parameter {k}")
            setattr(p, k, v)
        self.params = p
        self.n = int(n)
        dev = np.ascontiguousarray(devices if devices is not None else [p.device] * self.n, dtype=np.int32)
        self.g = C.c_void_p()
        rc = self.L.pp_group_create(C.byref(p), self.n, dev, C.byref(self.g))
        if rc != 0:
            msg = self.L.pp_group_last_error(self.g).decode() if self.g else "pp_group_create failed"
            if self.g:
                self.L.pp_group_destroy(self.g)
            self.g = None
            raise EngineError(msg)
        self.N = 0

    def _check(self, rc):
        if rc != 0:
            raise EngineError(self.L.pp_group_last_error(self.g).decode())

    def close(self):
        if getattr(self, "g", None):
            self.L.pp_group_destroy(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload(self, box, r, orient):
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.ascontiguousarray(orient, dtype=np.float64).reshape(-1, 9)
        self.N = int(r.shape[0])
        self._check(self.L.pp_group_upload(self.g, self.N, np.ascontiguousarray(box, dtype=np.float64), r, orient))

    def download(self):
        r = np.zeros((max(self.N, 1), 3))
        m = np.zeros((max(self.N, 1), 9))
        self._check(self.L.pp_group_download(self.g, r, m))
        return r[:self.N], m[:self.N].reshape(-1, 3, 3)

    def total_bonds(self):
        b, ov = C.c_longlong(0), C.c_int(0)
        self._check(self.L.pp_group_total_bonds(self.g, C.byref(b), C.byref(ov)))
        return int(b.value), int(ov.value)

    def bond_lists(self, max_per_particle=8):
        c = np.zeros(max(self.N, 1), dtype=np.int32)
        nb = np.full((max(self.N, 1), int(max_per_particle)), -1, dtype=np.int32)
        self._check(self.L.pp_group_bond_lists(self.g, int(max_per_particle), c, nb.reshape(-1)))
        return c[:self.N], nb[:self.N]

    def sweep(self, nsteps=1, sync=True):
        self._check(self.L.pp_group_sweep(self.g, int(nsteps)))
        if sync:
            self.sync()

    def sync(self):
        self._check(self.L.pp_group_sync(self.g))

    def observables(self, reset=False):
        o = Observables()
        self._check(self.L.pp_group_get_observables(self.g, C.byref(o), int(reset)))
        return o

    def check_energy(self):
        b, ov, ok = C.c_longlong(0), C.c_int(0), C.c_int(0)
        self._check(self.L.pp_group_check_energy(self.g, C.byref(b), C.byref(ov), C.byref(ok)))
        return int(b.value), int(ov.value), bool(ok.value)

    def volume_move(self, u_dir, u_vol, u_acc):
        a = C.c_int(0)
        self._check(self.L.pp_group_volume_move(self.g, u_dir, u_vol, u_acc, C.byref(a)))
        return int(a.value)

    def Lx_move(self, u_lx, u_acc):
        a = C.c_int(0)
        self._check(self.L.pp_group_Lx_move(self.g, u_lx, u_acc, C.byref(a)))
        return int(a.value)

    def slab_ranges(self):
        """[(x0, x1, nx, n_local)] of the slabs"""
        out = []
        for k in range(self.n):
            h = C.c_void_p()
            self._check(self.L.pp_group_engine(self.g, k, C.byref(h)))
            v = [C.c_int(0) for _ in range(4)]
            self.L.pp_slab_range(h, *[C.byref(x) for x in v])
            out.append(tuple(int(x.value) for x in v))
        return out
