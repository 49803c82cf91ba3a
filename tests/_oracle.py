"""ctypes bindings for the test-only oracle libraries (oracle/libpporacle.so, oracle/_ref/libppref.so).

TEST INFRASTRUCTURE: imported only from tests/, bench.py's cpu_baseline / --impl reference legs and
__graft_entry__.smoke().  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build_oracle():
    """(Re)build the restatement; builds _ref/ too when /root/reference is present."""
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True, stdout=subprocess.DEVNULL)


def have_ref():
    return os.path.exists(os.path.join(REF_DIR, "libppref.so"))


class SweepStats(C.Structure):
    _fields_ = [("tries", C.c_longlong), ("accepted", C.c_longlong), ("confined", C.c_longlong),
                ("overlapped", C.c_longlong), ("w_tries", C.c_uint64), ("w_accepted", C.c_uint64),
                ("pair_checks", C.c_longlong)]


_port = None


def port_lib():
    global _port
    if _port is not None:
        return _port
    path = os.path.join(ORACLE_DIR, "libpporacle.so")
    if not os.path.exists(path):
        build_oracle()
    L = C.CDLL(path)
    L.ppo_create.restype = C.c_void_p
    L.ppo_create.argtypes = [C.c_int, _dp, _dp, _dp, C.c_double, C.c_double, C.c_double]
    L.ppo_destroy.argtypes = [C.c_void_p]
    L.ppo_read_rrr.argtypes = [C.c_char_p, C.POINTER(C.c_longlong), C.POINTER(C.c_int), _dp,
                               C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.POINTER(C.c_double))]
    L.ppo_write_rrr.argtypes = [C.c_char_p, C.c_longlong, C.c_int, _dp, _dp, _dp]
    L.ppo_free.argtypes = [C.c_void_p]
    L.ppo_base_patches.argtypes = [_dp]
    L.ppo_orientation_from_rows.argtypes = [_dp, _dp, _dp]
    L.ppo_patches_from_orientation.argtypes = [_dp, _dp]
    L.ppo_gram_schmidt.argtypes = [_dp]
    L.ppo_rotate.argtypes = [_dp, _dp, C.c_double, _dp]
    L.ppo_axis_from_pair.argtypes = [C.c_double, C.c_double, _dp]
    L.ppo_cell_coords.argtypes = [_dp, _dp, _ip, _ip]
    L.ppo_pair.argtypes = [_dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double]
    L.ppo_N.argtypes = [C.c_void_p]
    L.ppo_get_box.argtypes = [C.c_void_p, _dp]
    L.ppo_get_cells.argtypes = [C.c_void_p, _ip]
    L.ppo_get_state.argtypes = [C.c_void_p, _dp, _dp]
    L.ppo_set_state.argtypes = [C.c_void_p, _dp, _dp, C.c_void_p]
    L.ppo_particle_bonds.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.ppo_total_bonds.restype = C.c_longlong
    L.ppo_total_bonds.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    L.ppo_bond_counts.argtypes = [C.c_void_p, _ip, _ip]
    L.ppo_bond_lists.argtypes = [C.c_void_p, C.c_int, _ip, _ip]
    L.ppo_eval_proposals.argtypes = [C.c_void_p, C.c_int, _ip, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip]
    L.ppo_volume_move.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                  C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    L.ppo_Lx_move.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                              C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
    L.ppo_neighbour_query.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, _ip]
    L.ppo_add_remove.argtypes = [C.c_void_p, C.c_double, C.c_int, C.c_int, _dp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                 C.POINTER(C.c_longlong)]
    L.ppo_sweep.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_double,
                            C.POINTER(C.c_longlong), C.POINTER(SweepStats)]
    L.ppo_even_grid.argtypes = [_dp, C.c_double, _ip]
    L.ppo_sweep_shift.argtypes = [C.c_uint64, C.c_uint64, _dp, _ip, _dp]
    L.ppo_philox4x32.argtypes = [C.c_uint32] * 6 + [np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
    _port = L
    return L


def read_rrr(path):
    """-> (step, box[3], r[N,3], orient[N,3,3]) parsed the way the reference parses it (system.c:213-245)."""
    L = port_lib()
    step = C.c_longlong(0)
    n = C.c_int(0)
    box = np.zeros(3)
    rp = C.POINTER(C.c_double)()
    mp = C.POINTER(C.c_double)()
    rc = L.ppo_read_rrr(path.encode(), C.byref(step), C.byref(n), box, C.byref(rp), C.byref(mp))
    if rc != 0:
        raise IOError(f"ppo_read_rrr({path}) failed with {rc}")
    N = n.value
    r = np.ctypeslib.as_array(rp, shape=(max(N, 1), 3))[:N].copy()
    m = np.ctypeslib.as_array(mp, shape=(max(N, 1), 3, 3))[:N].copy()
    L.ppo_free(rp)
    L.ppo_free(mp)
    return step.value, box, r, m


def write_rrr(path, step, box, r, m):
    L = port_lib()
    rc = L.ppo_write_rrr(path.encode(), step, r.shape[0], np.ascontiguousarray(box, dtype=np.float64),
                         np.ascontiguousarray(r, dtype=np.float64), np.ascontiguousarray(m, dtype=np.float64))
    if rc != 0:
        raise IOError(f"ppo_write_rrr({path}) failed with {rc}")


class PortSystem:
    """The CPU restatement (oracle/pp_oracle.c)."""

    def __init__(self, box, r, orient, kf_delta=0.119, kf_cosmax=0.98, temperature=0.1):
        self.L = port_lib()
        self.N = int(r.shape[0])
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.ascontiguousarray(orient, dtype=np.float64).reshape(-1, 9)
        box = np.ascontiguousarray(box, dtype=np.float64)
        if self.N == 0:
            r = np.zeros((1, 3))
            orient = np.zeros((1, 9))
        self.h = self.L.ppo_create(self.N, box, r, orient, kf_delta, kf_cosmax, temperature)

    def close(self):
        if self.h:
            self.L.ppo_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def box(self):
        b = np.zeros(3)
        self.L.ppo_get_box(self.h, b)
        return b

    def cells(self):
        c = np.zeros(3, dtype=np.int32)
        self.L.ppo_get_cells(self.h, c)
        return c

    def state(self):
        r = np.zeros((max(self.N, 1), 3))
        m = np.zeros((max(self.N, 1), 9))
        self.L.ppo_get_state(self.h, r, m)
        return r[:self.N], m[:self.N].reshape(-1, 3, 3)

    def set_state(self, r, orient, box=None):
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.ascontiguousarray(orient, dtype=np.float64).reshape(-1, 9)
        bp = None
        if box is not None:
            self._box_keep = np.ascontiguousarray(box, dtype=np.float64)
            bp = self._box_keep.ctypes.data_as(C.c_void_p)
        self.L.ppo_set_state(self.h, r, orient, bp)

    def total_bonds(self):
        ov = C.c_int(0)
        b = self.L.ppo_total_bonds(self.h, C.byref(ov))
        return int(b), int(ov.value)

    def bond_counts(self):
        c = np.zeros(max(self.N, 1), dtype=np.int32)
        o = np.zeros(max(self.N, 1), dtype=np.int32)
        self.L.ppo_bond_counts(self.h, c, o)
        return c[:self.N], o[:self.N]

    def bond_lists(self, max_nb=8):
        c = np.zeros(max(self.N, 1), dtype=np.int32)
        nb = np.full((max(self.N, 1), max_nb), -1, dtype=np.int32)
        self.L.ppo_bond_lists(self.h, max_nb, c, nb.reshape(-1))
        return c[:self.N], nb[:self.N]

    def eval_proposals(self, idx, disp, orient_new, u, commit=False):
        n = len(idx)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        disp = np.ascontiguousarray(disp, dtype=np.float64).reshape(-1, 3)
        orient_new = np.ascontiguousarray(orient_new, dtype=np.float64).reshape(-1, 9)
        u = np.ascontiguousarray(u, dtype=np.float64)
        dE = np.zeros(n, dtype=np.int32)
        ov = np.zeros(n, dtype=np.int32)
        acc = np.zeros(n, dtype=np.int32)
        self.L.ppo_eval_proposals(self.h, n, idx, disp, orient_new, u, int(commit), dE, ov, acc)
        return dE, ov, acc

    def volume_move(self, P, rescale_factor_max, u_dir, u_vol, u_acc, bonds):
        b = C.c_longlong(bonds)
        ov = C.c_int(0)
        acc = self.L.ppo_volume_move(self.h, P, rescale_factor_max, u_dir, u_vol, u_acc, C.byref(b), C.byref(ov))
        return int(acc), int(b.value), int(ov.value)

    def Lx_move(self, Lx_change_max, Lyz_min, Lyz_max, u_lx, u_acc, bonds, faithful_sign=True):
        b = C.c_longlong(bonds)
        ov = C.c_int(0)
        acc = self.L.ppo_Lx_move(self.h, Lx_change_max, Lyz_min, Lyz_max, int(faithful_sign), u_lx, u_acc,
                                 C.byref(b), C.byref(ov))
        return int(acc), int(b.value), int(ov.value)

    def patch_neighbours(self, i):
        """_set_neighbours (avb.c:43-75) -> (neighbour per patch, number of bonded neighbours)"""
        nb = np.full(4, -1, dtype=np.int32)
        n = self.L.ppo_neighbour_query(self.h, int(i), 0, 4, nb)
        return nb, int(n)

    def range_neighbours(self, i, max_nb=32):
        nb = np.full(max_nb, -1, dtype=np.int32)
        n = self.L.ppo_neighbour_query(self.h, int(i), 1, int(max_nb), nb)
        return int(n), nb

    def add_remove(self, z, n_min, n_max, uniforms, bonds):
        """MC_add_remove (MC.c:342-448) with recorded uniforms -> (kind, accepted, uniforms consumed, N, running bonds)"""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        a, used, b = C.c_int(0), C.c_int(0), C.c_longlong(bonds)
        kind = self.L.ppo_add_remove(self.h, float(z), int(n_min), int(n_max), u, int(u.size), C.byref(a), C.byref(used), C.byref(b))
        self.N = int(self.L.ppo_N(self.h))
        return int(kind), int(a.value), int(used.value), self.N, int(b.value)

    def sweep(self, seed, sweep_index, disp_max, theta_max, bonds, stats=None, shift_every=4):
        b = C.c_longlong(bonds)
        st = stats if stats is not None else SweepStats()
        self.L.ppo_sweep(self.h, seed, sweep_index, shift_every, disp_max, theta_max, C.byref(b), C.byref(st))
        return int(b.value), st


def sorted_bond_lists(nb):
    """Canonical form of a neighbour table for set comparison: every row ascending, the -1 padding last."""
    a = np.where(nb < 0, np.iinfo(np.int32).max, nb)
    a = np.sort(a, axis=1)
    return np.where(a == np.iinfo(np.int32).max, -1, a).astype(np.int32)


def rotate_port(m_old, axis, theta):
    out = np.zeros(9)
    port_lib().ppo_rotate(np.ascontiguousarray(m_old, dtype=np.float64).reshape(9),
                          np.ascontiguousarray(axis, dtype=np.float64), float(theta), out)
    return out.reshape(3, 3)


# --------------------------------------------------------------------------------------------
# The unmodified reference behind ref_harness.c.  One global System per process (the reference
# keeps file-scope state), so RefSystem is a singleton-style wrapper.
# --------------------------------------------------------------------------------------------
_ref = None


def ref_lib():
    global _ref
    if _ref is not None:
        return _ref
    path = os.path.join(REF_DIR, "libppref.so")
    L = C.CDLL(path)
    L.ref_open.argtypes = [C.c_char_p]
    L.ref_box.argtypes = [_dp]
    L.ref_cells.argtypes = [_ip]
    L.ref_running_energy.restype = C.c_double
    L.ref_set_running_energy.argtypes = [C.c_double]
    L.ref_counters.argtypes = [_ip, _ip]
    L.ref_get_state.argtypes = [_dp, _dp]
    L.ref_get_patches.argtypes = [_dp]
    L.ref_set_state.argtypes = [_dp, _dp, C.c_void_p]
    L.ref_total_energy.restype = C.c_double
    L.ref_total_energy.argtypes = [C.POINTER(C.c_int)]
    L.ref_bond_counts.argtypes = [_ip, _ip]
    L.ref_bond_lists.argtypes = [C.c_int, _ip, _ip]
    L.ref_rotate.argtypes = [_dp, _dp, C.c_double, _dp]
    L.ref_patch_pairs.argtypes = [C.c_int, _ip, _ip, _ip, _ip]
    L.ref_eval_proposals.argtypes = [C.c_int, _ip, _dp, _dp, _dp, C.c_int, _ip, _ip, _ip]
    L.ref_run_moves.argtypes = [C.c_long]
    L.ref_run_steps.argtypes = [C.c_long]
    L.ref_set_stream.argtypes = [C.c_void_p, C.c_long]
    L.ref_stream_pos.restype = C.c_long
    L.ref_draws.restype = C.c_long
    L.ref_seed.argtypes = [C.c_long]
    L.ref_random_vector_on_sphere.argtypes = [_dp]
    L.ref_gram_schmidt.argtypes = [_dp]
    L.ref_volume.restype = C.c_double
    _ref = L
    return L


def write_input(path, rrr, **keys):
    """Writes a reference-style input file (key = value lines, README.md:18-80)."""
    base = dict(Dynamics=0, Ensemble=0, Temperature=0.1, Steps=0, Disp_max=0.1, Theta_max=0.1,
                Initial_conditions_file=rrr, Print_every=1000000, Save_every=1000000,
                Configuration_folder=".", KF_delta=0.119, KF_cosmax=0.98, Seed=12345)
    base.update(keys)
    with open(path, "w") as f:
        for k, v in base.items():
            f.write(f"{k} = {v}\n")


class RefSystem:
    """Drives the unmodified reference through oracle/ref_harness.c."""

    def __init__(self, rrr, **keys):
        self.L = ref_lib()
        self.tmp = tempfile.mkdtemp(prefix="ppref_")
        inp = os.path.join(self.tmp, "input")
        write_input(inp, rrr, **keys)
        devnull = os.open(os.devnull, os.O_WRONLY)
        saved = os.dup(2)
        os.dup2(devnull, 2)  # the reference logs INFO lines to stderr for every optional key
        try:
            rc = self.L.ref_open(inp.encode())
        finally:
            os.dup2(saved, 2)
            os.close(devnull)
            os.close(saved)
        if rc != 0:
            raise RuntimeError(f"ref_open failed: {rc}")
        self.N = self.L.ref_N()

    def close(self):
        self.L.ref_close()

    def box(self):
        b = np.zeros(3)
        self.L.ref_box(b)
        return b

    def cells(self):
        c = np.zeros(3, dtype=np.int32)
        self.L.ref_cells(c)
        return c

    def state(self):
        r = np.zeros((self.N, 3))
        m = np.zeros((self.N, 9))
        self.L.ref_get_state(r, m)
        return r, m.reshape(-1, 3, 3)

    def patches(self):
        p = np.zeros((self.N, 12))
        self.L.ref_get_patches(p)
        return p.reshape(-1, 4, 3)

    def set_state(self, r, orient, box=None):
        r = np.ascontiguousarray(r, dtype=np.float64).reshape(-1, 3)
        orient = np.ascontiguousarray(orient, dtype=np.float64).reshape(-1, 9)
        bp = None
        if box is not None:
            self._box_keep = np.ascontiguousarray(box, dtype=np.float64)
            bp = self._box_keep.ctypes.data_as(C.c_void_p)
        self.L.ref_set_state(r, orient, bp)

    def total_energy(self):
        ov = C.c_int(0)
        e = self.L.ref_total_energy(C.byref(ov))
        return float(e), int(ov.value)

    def bond_counts(self):
        c = np.zeros(self.N, dtype=np.int32)
        o = np.zeros(self.N, dtype=np.int32)
        self.L.ref_bond_counts(c, o)
        return c, o

    def bond_lists(self, max_nb=8):
        """Bonded neighbours by the reference's MC_interact over all pairs (ascending index)."""
        c = np.zeros(max(self.N, 1), dtype=np.int32)
        nb = np.full((max(self.N, 1), max_nb), -1, dtype=np.int32)
        self.L.ref_bond_lists(max_nb, c, nb.reshape(-1))
        return c[:self.N], nb[:self.N]

    def patch_pairs(self, max_nb=8):
        """Bonded neighbours with MC_interact's patch indices (onp, onq), over all pairs in ascending index."""
        n = max(self.N, 1)
        c = np.zeros(n, dtype=np.int32)
        q, onp, onq = (np.full((n, max_nb), -1, dtype=np.int32) for _ in range(3))
        self.L.ref_patch_pairs(max_nb, c, q.reshape(-1), onp.reshape(-1), onq.reshape(-1))
        return c[:self.N], q[:self.N], onp[:self.N], onq[:self.N]

    def rotate(self, m_old, axis, theta):
        out = np.zeros(9)
        self.L.ref_rotate(np.ascontiguousarray(m_old, dtype=np.float64).reshape(9),
                          np.ascontiguousarray(axis, dtype=np.float64), float(theta), out)
        return out.reshape(3, 3)

    def eval_proposals(self, idx, disp, orient_new, u, commit=False):
        n = len(idx)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        disp = np.ascontiguousarray(disp, dtype=np.float64).reshape(-1, 3)
        orient_new = np.ascontiguousarray(orient_new, dtype=np.float64).reshape(-1, 9)
        u = np.ascontiguousarray(u, dtype=np.float64)
        dE = np.zeros(n, dtype=np.int32)
        ov = np.zeros(n, dtype=np.int32)
        acc = np.zeros(n, dtype=np.int32)
        self.L.ref_eval_proposals(n, idx, disp, orient_new, u, int(commit), dE, ov, acc)
        return dE, ov, acc

    def set_stream(self, stream):
        if stream is None:
            self.L.ref_set_stream(None, 0)
            self._stream = None
        else:
            self._stream = np.ascontiguousarray(stream, dtype=np.float64)
            self.L.ref_set_stream(self._stream.ctypes.data_as(C.c_void_p), len(self._stream))

    def running_energy(self):
        return float(self.L.ref_running_energy())

    def set_running_energy(self, e):
        self.L.ref_set_running_energy(float(e))

    def counters(self):
        t = np.zeros(8, dtype=np.int32)
        a = np.zeros(8, dtype=np.int32)
        self.L.ref_counters(t, a)
        return t, a

    def run_steps(self, n):
        """n calls of do_ensemble (main.c:119), drawing from the harness' POSIX drand48"""
        self.L.ref_run_steps(int(n))

    def run_moves(self, n):
        self.L.ref_run_moves(int(n))

    def check_energy(self):
        self.L.ref_check_energy()

    def box_move(self, kind, uniforms):
        """One MC_change_volume (kind 'volume', MC.c:450-513) or MC_change_Lx (kind 'Lx', MC.c:515-585) fed with the
        given uniforms in the reference's draw order -> (accepted, uniforms consumed, box, energy, cells)."""
        slot = 6 if kind == "volume" else 7
        before = int(self.counters()[1][slot])
        self.set_stream(uniforms)
        if kind == "volume":
            self.L.ref_volume_move()
        else:
            self.L.ref_Lx_move()
        used = int(self.L.ref_stream_pos())
        self.set_stream(None)
        return int(self.counters()[1][slot]) - before, used, self.box(), self.running_energy(), self.cells()

    def add_remove(self, uniforms):
        """One MC_add_remove (MC.c:342-448) fed with the given uniforms in the reference's draw order ->
        (kind: 2 insertion, 3 removal, -1 returned before a try; accepted; uniforms consumed; N; running energy)."""
        t0, a0 = self.counters()
        self.set_stream(uniforms)
        self.L.ref_add_remove()
        used = int(self.L.ref_stream_pos())
        self.set_stream(None)
        t1, a1 = self.counters()
        kind = 2 if t1[2] > t0[2] else (3 if t1[3] > t0[3] else -1)
        accepted = int(a1[2] - a0[2] + a1[3] - a0[3])
        self.N = self.L.ref_N()
        return kind, accepted, used, self.N, self.running_energy()

    def random_vector_on_sphere(self, uniforms):
        """utils.c:115-131 fed with the given uniforms -> (axis, uniforms consumed)"""
        out = np.zeros(3)
        self.set_stream(uniforms)
        self.L.ref_random_vector_on_sphere(out)
        used = int(self.L.ref_stream_pos())
        self.set_stream(None)
        return out, used

    def gram_schmidt(self, m):
        out = np.ascontiguousarray(m, dtype=np.float64).reshape(9).copy()
        self.L.ref_gram_schmidt(out)
        return out.reshape(3, 3)


def run_generator(N, rho, out_path):
    """Runs the reference's own generator binary (deterministic: it never seeds drand48)."""
    gen = os.path.join(REF_DIR, "generator")
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([gen, str(N), str(rho)], cwd=d, check=True, stderr=subprocess.DEVNULL)
        os.replace(os.path.join(d, "generated.rrr"), out_path)
