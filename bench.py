#!/usr/bin/env python
"""bench.py -- attempted roto-translation moves per second of the checkerboard MC engine.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference CPU arm)

A "step" is one MC step of the reference (main.c:119): N_particles attempted roto-translations
(one 8-colour checkerboard sweep including its share of the cell builds).  The headline workload is the
same at every GPU count -- BASELINE.json configs[3], N = 4 194 304 at rho = 0.6, NVT -- on one engine at
--gpus 1 and split into N x-slabs at --gpus N (strong scaling), so the driver's efficiency divides
like by like (at --gpus N the cell grid has a multiple of 2 N layers along x, pp_params.x_layers_multiple, so that every
slab holds the same number of layers).  The one-GPU line also carries, under "configs", the other BASELINE configurations
measured in the same run: N = 65536 (configs[1]), NPT and Lx moves at N = 32768 (configs[2]), the FCC
crystal at N = 256 000 (configs[4]) and N = 1000 (configs[0]); --workload picks another headline,
--no-configs drops the extra ones.

Prints ONE JSON line on rank 0 (see the task contract): value = device-timed throughput with the
state resident in HBM (L2 flushed between steps), e2e = the same metric through the C ABI with
host buffers (H2D of the configuration + sweep + D2H of the observables every step; e2e_roundtrip
adds pp_download of the whole state), roofline for the sweep kernel from live CUDA-event timings,
cpu_baseline = the unmodified reference timed on this box's host cores on the same configuration.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

KF_DELTA, KF_COSMAX, DISP_MAX, THETA_MAX, SEED = 0.119, 0.98, 0.1, 0.1, 12345
WORKLOADS = {
    # kind "fluid": generator-made fluid; "fcc": the dofccrrr.py crystal; extra = engine / input-file keys on top of NVT
    "4M": dict(N=4194304, rho=0.6, T=0.1, kind="fluid", extra={}),
    "64k": dict(N=65536, rho=0.5, T=0.1, kind="fluid", extra={}),
    "1k": dict(N=1000, rho=0.1, T=0.1, kind="fluid", extra={}),
    "256k": dict(N=262144, rho=0.5, T=0.1, kind="fluid", extra={}),
    "512k": dict(N=524288, rho=0.6, T=0.1, kind="fluid", extra={}),
    "1M": dict(N=1048576, rho=0.6, T=0.1, kind="fluid", extra={}),
    "128k": dict(N=131072, rho=0.6, T=0.1, kind="fluid", extra={}),
    "448k": dict(N=458752, rho=0.6, T=0.1, kind="fluid", extra={}),
    "576k": dict(N=589824, rho=0.6, T=0.1, kind="fluid", extra={}),
    "640k": dict(N=655360, rho=0.6, T=0.1, kind="fluid", extra={}),
    "768k": dict(N=786432, rho=0.6, T=0.1, kind="fluid", extra={}),
    "npt32k": dict(N=32768, rho=0.3, T=0.15, kind="fluid", extra=dict(ensemble=5, P=0.5, rescale_factor_max=1e-3)),
    "lx32k": dict(N=32768, rho=0.3, T=0.15, kind="fluid", extra=dict(Lx_move=1, Lx_change_max=0.01, Lyz_min=20., Lyz_max=80.)),
    "fcc256k": dict(N=256000, rho=1.041, T=0.11, kind="fcc", extra={}),
}
DEFAULT_WORKLOAD = "4M"
EXTRA_CONFIGS = ["64k", "npt32k", "lx32k", "fcc256k", "1k"]
MELT_SWEEPS = 300  # hard-sphere sweeps that turn the thinned lattice of patchy_generate into a fluid
# dram__bytes_read.sum + dram__bytes_write.sum per colour-pass launch from the committed ncu captures (profiles/), by
# (workload, GPUs); absent where no capture of that shape is committed
SWEEP_TRAFFIC = {
    ("4M", 1): 587.4e6,    # profiles/r03u_launches_4M.txt: k_sweep_fused 541.3 MB read + 46.1 MB written per launch (524288 trials)
    ("512k", 1): 39.0e6,   # profiles/r02n_sweep_fused_512k_full.txt: 38.7 MB read + 0.3 MB written per launch (65536 trials), state resident in L2
    ("fcc256k", 1): 37.2e6,  # profiles/r02y_fcc_and_slab_upload.txt (capture of the fused launch on the crystal; the tile kernel now runs it)
}
SWEEP_BYTES_PER_ATTEMPT = 64   # state read once + written once per sweep at S = 32 B
CELLS_BYTES_PER_PARTICLE = 72
BONDS_BYTES_PER_PARTICLE = 32
RESCALE_BYTES_PER_PARTICLE = 32
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
CACHE = os.path.join(tempfile.gettempdir(), "pp_bench_cache")


def workload_name(w):
    d = WORKLOADS[w]
    ens = "NPT P=%g rescale_factor_max=%g" % (d["extra"]["P"], d["extra"]["rescale_factor_max"]) if d["extra"].get("ensemble") == 5 else "NVT"
    if d["extra"].get("Lx_move"):
        ens += " Lx_move=1 Lx_change_max=%g" % d["extra"]["Lx_change_max"]
    what = "FCC crystal 40x40x40x4, " if d["kind"] == "fcc" else ""
    return f"{what}N={d['N']} rho={d['rho']} {ens} T={d['T']} KF_delta={KF_DELTA} KF_cosmax={KF_COSMAX} Disp_max=Theta_max={DISP_MAX}"


# ------------------------------------------------------------------------------------------------
# configurations: made once per box, cached as files both arms read
# ------------------------------------------------------------------------------------------------
def _reference_generator(N, rho):
    """generated.rrr of the reference's own generator (generator.c never seeds drand48: deterministic)."""
    gen = os.path.join(REF_DIR, "generator")
    if not os.path.exists(gen) or N > 300000:
        return None
    out = os.path.join(CACHE, f"refgen_{N}_{rho}.rrr")
    if not os.path.exists(out):
        d = tempfile.mkdtemp(prefix="ppgen_")
        try:
            subprocess.run([gen, str(N), str(rho)], cwd=d, check=True, stderr=subprocess.DEVNULL, stdout=subprocess.DEVNULL)
            shutil.move(os.path.join(d, "generated.rrr"), out)
        except Exception:
            return None
        finally:
            shutil.rmtree(d, ignore_errors=True)
    return out


def _own_generator(N, rho, device):
    """host/patchy_generate: thinned lattice melted by MELT_SWEEPS hard-sphere sweeps (a separate process; binary .rrb)."""
    gen = os.path.join(ROOT, "host", "patchy_generate")
    out = os.path.join(CACHE, f"gen_{N}_{rho}_melt{MELT_SWEEPS}.rrb")
    desc = f"host/patchy_generate {N} {rho} seed 1, melted by {MELT_SWEEPS} hard-sphere sweeps"
    if os.path.exists(out):
        return out, desc
    if os.path.exists(gen):
        tmp = out + f".tmp{os.getpid()}.rrb"
        p = subprocess.run([gen, str(N), str(rho), "1", tmp, str(MELT_SWEEPS), str(device)], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if p.returncode == 0 and os.path.exists(tmp):
            os.replace(tmp, out)
            return out, desc
        if os.path.exists(tmp):
            os.remove(tmp)
    return None, None


def make_configuration(w, device=0):
    """-> box, r, m (N x 3 x 3), description"""
    import patchyparticles_b200 as pkg
    os.makedirs(CACHE, exist_ok=True)
    d = WORKLOADS[w]
    N, rho = d["N"], d["rho"]
    if d["kind"] == "fcc":
        box, r, m = pkg.fcc_configuration(rho, 40, 40, 40)
        return box, r, m, "FCC lattice of dofccrrr.py, 40 x 40 x 40 cells of 4"
    path = _reference_generator(N, rho)
    if path is not None:
        _, box, r, m = pkg.read_rrr(path)
        return box, r, m, f"reference generator {N} {rho}"
    path, desc = _own_generator(N, rho, device)
    if path is not None:
        _, box, r, m = pkg.read_rrr(path)
        return box, r, m, desc
    box, r, m = pkg.random_fluid(N, rho, seed=1)
    return box, r, m, f"seeded lattice-thinned hard-sphere configuration {N} {rho} (no generator binary found)"


def reference_text_file(w, device=0):
    """The workload's configuration as the text .rrr the reference reads (system.c:198-252)."""
    d = WORKLOADS[w]
    os.makedirs(CACHE, exist_ok=True)
    if d["kind"] == "fluid":
        path = _reference_generator(d["N"], d["rho"])
        if path is not None:
            return path
    out = os.path.join(CACHE, f"text_{w}.rrr")
    if not os.path.exists(out):
        import _oracle  # test-only C writer (output.c:302-317 format); allowed in the reference / cpu_baseline legs
        box, r, m, _ = make_configuration(w, device)
        _oracle.write_rrr(out + ".tmp", 0, box, r, m.reshape(-1, 9))
        os.replace(out + ".tmp", out)
    return out


# ------------------------------------------------------------------------------------------------
# reference CPU arm
# ------------------------------------------------------------------------------------------------
def _reference_binaries():
    """The north-star flavour (-O3 -march=native) compiled on THIS host when gcc and the reference sources are here
    (the build container), then the prebuilt ones (the GPU box has neither the sources nor, possibly, the same CPU)."""
    out = []
    src = "/root/reference"
    if os.path.isdir(src) and shutil.which("gcc"):
        exe = os.path.join(CACHE, "PatchyParticles_native_here")
        if not os.path.exists(exe):
            os.makedirs(CACHE, exist_ok=True)
            files = [os.path.join(src, f) for f in ("main.c", "MC.c", "avb.c", "vmmc.c", "cells.c", "output.c", "parse_input.c", "system.c", "utils.c")]
            if all(os.path.exists(f) for f in files):
                subprocess.run(["gcc", "-O3", "-march=native", "-o", exe] + files + ["-lm"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if os.path.exists(exe):
            out.append((exe, "-O3 -march=native, compiled on this host"))
    out.append((os.path.join(REF_DIR, "PatchyParticles_native"), "-O3 -march=native of the build host"))
    out.append((os.path.join(REF_DIR, "PatchyParticles"), "-O3 -march=x86-64-v3"))
    return [(b, f) for b, f in out if os.path.exists(b)]


def _run_ref(binary, workdir, rrr, steps, w):
    d = WORKLOADS[w]
    x = d["extra"]
    keys = (f"Dynamics = 0\nEnsemble = {x.get('ensemble', 0)}\nTemperature = {d['T']}\nSteps = {steps}\nDisp_max = {DISP_MAX}\n"
            f"Theta_max = {THETA_MAX}\nInitial_conditions_file = {rrr}\nPrint_every = 1000000000\n"
            f"Save_every = 1000000000\nConfiguration_folder = .\nKF_delta = {KF_DELTA}\nKF_cosmax = {KF_COSMAX}\n"
            f"Seed = {SEED}\nLog_file = none\n")
    if x.get("ensemble") == 5:
        keys += f"P = {x['P']}\nrescale_factor_max = {x['rescale_factor_max']}\n"
    if x.get("Lx_move"):
        keys += f"Lx_move = 1\nLx_change_max = {x['Lx_change_max']}\nLyz_min = {x['Lyz_min']}\nLyz_max = {x['Lyz_max']}\n"
    with open(os.path.join(workdir, "input"), "w") as f:
        f.write(keys)
    t0 = time.perf_counter()
    p = subprocess.run([binary, "input"], cwd=workdir, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, p.returncode


def time_reference(w, sweeps, warmup=0, device=0):
    """wall(sweeps) - wall(0 sweeps) of the unmodified reference binary, single thread (the reference
    has no threads: main.c:26-143); `warmup` steps are run once before, untimed (page cache, CPU clocks)."""
    N = WORKLOADS[w]["N"]
    rrr = reference_text_file(w, device)
    d = tempfile.mkdtemp(prefix="ppref_")
    try:
        chosen = None
        for binary, flags in _reference_binaries():
            t_init, rc = _run_ref(binary, d, rrr, 0, w)
            if rc == 0:
                chosen = (binary, flags)
                break
        if chosen is None:
            return None
        binary, flags = chosen
        if warmup > 0:
            _run_ref(binary, d, rrr, warmup, w)
        t_run, rc = _run_ref(binary, d, rrr, sweeps, w)
        if rc != 0:
            return None
        dt = max(t_run - t_init, 1e-9)
        return {"value": sweeps * N / dt, "seconds": dt, "init_seconds": t_init, "flags": flags, "sweeps": sweeps}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_sample_sweeps(w, budget_s=15.0):
    N = WORKLOADS[w]["N"]
    est = N / (4.0e5 if N < 500000 else 1.5e5)  # attempts/s/core measured at survey time (cache-bound at 4M)
    return max(1, int(budget_s / est))


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = WORKLOADS[w]["N"]
    sweeps = max(1, min(args.steps, reference_sample_sweeps(w, 40.0)))
    warm = max(0, min(args.warmup, 3, sweeps)) if N < 500000 else 0  # a separate, untimed run of the binary
    res = time_reference(w, sweeps, warm)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref reference binaries missing or not runnable on this host"}))
        return 0
    line = {
        "impl": "reference", "metric": "MC roto-translation attempts/sec", "value": res["value"], "unit": "attempts/s",
        "n_gpus": args.gpus, "steps": sweeps, "warmup": warm, "ms_per_step": 1e3 * res["seconds"] / sweeps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(w), "N_total": N},
        "cpu_baseline": {"value": res["value"], "unit": "attempts/s", "cores": 1, "kind": "reference",
                         "sample": f"{sweeps} MC steps of the unmodified reference ({res['flags']}) on the same configuration file as the GPU arm, "
                                   "wall(steps) - wall(0 steps); the reference is single-threaded"},
        "e2e": {"value": res["value"], "unit": "attempts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark(self):
        return len(self.rows)

    def stop(self, first=0):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows[first:]:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


_FLUSH = {}


def flush_buffer(dev):
    import torch
    if dev not in _FLUSH:
        _FLUSH[dev] = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    return _FLUSH[dev]


class Arm:
    """The engine under test: one Engine or one SlabEngine rank."""

    def __init__(self, pkg, args, w, rank, local_rank, world, sharded):
        import torch
        d = WORKLOADS[w]
        params = dict(temperature=d["T"], kf_delta=KF_DELTA, kf_cosmax=KF_COSMAX, disp_max=DISP_MAX, theta_max=THETA_MAX,
                      lanes_per_cell=args.lanes, cells_max_occupancy=args.cells_max_occupancy, cell_lanes=args.cell_lanes,
                      cells_mid_occupancy=args.cells_mid_occupancy)
        params.update(d["extra"])
        self.sharded = sharded
        if sharded:
            from patchyparticles_b200.slab import SlabEngine
            # every slab the same number of x layers (170 -> 168 layers on 2 and 4 GPUs, 160 on 8): no engine carries more than its share, and
            # the fused pass of an engine stays inside one resident wave of blocks (profiles/r03h_one_wave.txt)
            params["x_layers_multiple"] = 2 * world
            self.drv = SlabEngine(rank, world, local_rank, seed=SEED, **params)
            self.eng = self.drv.eng
        else:
            self.eng = pkg.Engine(seed=SEED + rank, device=local_rank, **params)
            self.eng.set_stream(torch.cuda.current_stream().cuda_stream)
            self.drv = self.eng

    def upload(self, box, r, m):
        self.drv.upload(box, r, m)

    def total_bonds(self):
        return self.drv.total_bonds()

    def sweep(self, n, sync=False):
        self.drv.sweep(n, sync=sync)

    def observables(self):
        o = self.drv.observables()
        if self.sharded:
            return o
        return {"pair_checks": o.pair_checks, "tries": list(o.tries), "accepted": list(o.accepted), "box": list(o.box), "energy": o.energy}

    def check_energy(self):
        return self.drv.check_energy()

    def download(self, out=None):
        if out is not None and self.drv is self.eng:
            return self.drv.download(out)
        return self.drv.download()

    def cells(self):
        return list(self.eng.observables().n_side)

    def close(self):
        self.drv.close()


def measure(pkg, args, w, K, W, n_e2e, cpu_budget_s, ctx, headline):
    """One workload on the current process group -> the fields of a bench line (rank 0) or None."""
    import numpy as np
    import torch
    rank, local_rank, world, dev, barrier, dist = ctx
    sharded = world > 1 and not args.replicas
    d = WORKLOADS[w]
    N = d["N"]
    box_moves = d["extra"].get("ensemble") == 5 or bool(d["extra"].get("Lx_move"))
    if world > 1:  # rank 0 makes the cached files, everybody reads them
        if rank == 0:
            make_configuration(w, local_rank)
        barrier()
    box, r, m, data_desc = make_configuration(w, local_rank)
    stream = torch.cuda.current_stream()
    arm = Arm(pkg, args, w, rank, local_rank, world, sharded)
    eng = arm.eng
    arm.upload(box, r, m)
    bonds0, ov0 = arm.total_bonds()
    assert ov0 == 0, "initial configuration overlaps"
    n_systems = 1 if sharded else world          # independent systems in flight
    attempts_per_step = float(N) * n_systems      # whole-job attempts per step
    flush = flush_buffer(dev)

    # ---- warm-up ----------------------------------------------------------------------------
    arm.sweep(W, sync=True)
    cells = arm.cells()
    flush.zero_()
    barrier()

    # ---- timed: K steps, L2 flushed between steps, per-step CUDA events on the launch stream ----
    sampler = ClockSampler(local_rank) if (rank == 0 and headline) else None
    if sampler:
        sampler.start()
        time.sleep(0.1)
    eng.reset_kernel_stats()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    o0 = arm.observables()
    barrier()
    mark = sampler.mark() if sampler else 0
    t_wall0 = time.perf_counter()
    for i in range(K):
        flush.zero_()
        starts[i].record(stream)
        arm.sweep(1)
        ends[i].record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = np.array([s_.elapsed_time(e_) for s_, e_ in zip(starts, ends)])
    launches_timed = sum(v[0] for v in eng.kernel_stats().values())
    total_ms = float(step_ms.sum())
    o1 = arm.observables()

    # ---- the same steps back to back without the flush (steady state of a real run) -----------------
    k_warm = max(K, 50)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    arm.sweep(k_warm)
    ev1.record(stream)
    barrier()
    warm_ms = ev0.elapsed_time(ev1)
    o2 = arm.observables()
    pair_checks_per_s = (o2["pair_checks"] - o1["pair_checks"]) * (1 if sharded else world) / (warm_ms * 1e-3)
    clocks = sampler.stop(mark) if sampler else None

    # ---- per-kernel timing for the roofline (CUDA events around each kernel class) -----------------
    eng.set_profiling(True)
    eng.reset_kernel_stats()
    n_prof = min(K, 300)
    for i in range(n_prof):
        flush.zero_()
        arm.sweep(1)
    if not box_moves:
        arm.total_bonds()
    eng.sync()
    stats = eng.kernel_stats()
    eng.set_profiling(False)
    n_local = eng.slab_range()[3] if sharded else N
    peak, peak_src = measured_peaks()
    sweep_launches, sweep_ms = stats["sweep"]
    n_passes = 8 * n_prof
    passes_per_launch = max(int(round(n_passes / max(sweep_launches, 1))), 1)  # 1: a launch per colour pass; 4: a dataflow launch per half sweep
    per_launch_ms = sweep_ms / max(sweep_launches, 1)
    dense = args.cells_max_occupancy < 0 and args.lanes != 1 and 27. * N > 32. * cells[0] * cells[1] * cells[2]  # routing_limits (pp_api.cu)
    tile_only = (n_local < 100000 and args.cells_max_occupancy < 0 and args.lanes != 1) or min(cells) < 4 or dense
    sweep_kernel = ("k_sweep_pass (tile kernel)" if tile_only else
                    "k_sweep_half (the four colour passes of a half sweep in one dataflow launch; 8 / 4 / 1 lanes per cell)" if passes_per_launch == 4 else
                    "k_sweep_fused (8 / 4 / 1 lanes per cell in one launch)")
    n_owned = N / world if sharded else N
    bytes_per_launch = SWEEP_BYTES_PER_ATTEMPT * n_owned * passes_per_launch / 8.0
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    cell_builds = max(stats["cells"][0] // 10, 1)
    n_bonds = max(stats["bonds"][0], 1)
    # a full count is three launches (zero, FP32 screen, FP64 on the queued pairs) on a grid with four or more cells per axis, else two
    per_count = 3 if min(cells) >= 4 else 2
    ms_per_count = stats["bonds"][1] / max(stats["bonds"][0] / per_count, 1.)
    kernels = {
        "sweep_pass": {"launches": sweep_launches, "passes": n_passes, "ms_per_launch": per_launch_ms, "ms_per_pass": sweep_ms / n_passes,
                       "alg_bytes_per_launch": bytes_per_launch, "GBps": achieved},
        "cell_build": {"launches": stats["cells"][0], "builds": cell_builds, "ms_per_build": stats["cells"][1] / cell_builds,
                       "GBps": CELLS_BYTES_PER_PARTICLE * n_local / (max(stats["cells"][1], 1e-9) / cell_builds * 1e-3) / 1e9},
        "proposals": {"launches": stats["rng"][0], "ms_per_launch": stats["rng"][1] / max(stats["rng"][0], 1)},
        "bond_count": {"launches": stats["bonds"][0], "ms_per_launch": stats["bonds"][1] / n_bonds, "ms_per_count": ms_per_count,
                       "GBps": BONDS_BYTES_PER_PARTICLE * n_owned / (max(ms_per_count, 1e-9) * 1e-3) / 1e9},
    }
    if stats["rescale"][0]:
        n_rs = stats["rescale"][0]
        kernels["rescale"] = {"launches": n_rs, "ms_per_launch": stats["rescale"][1] / n_rs,
                              "GBps": RESCALE_BYTES_PER_PARTICLE * n_owned / (max(stats["rescale"][1] / n_rs, 1e-9) * 1e-3) / 1e9}
    if sharded:  # phase waits / signals between neighbouring engines, record pushes and pulls at a grid shift
        kernels["halo"] = {"launches": stats["halo"][0], "ms_per_sweep": stats["halo"][1] / max(n_prof, 1)}

    # invariant after the measured regions: running energy == recount (MC_check_energy)
    b_chk, ov_chk, ok = arm.check_energy()
    assert ok, "energy check failed after the timed region"

    # ---- e2e through the public API with host buffers ---------------------------------------------
    # the step's input is the configuration the run has reached (pinned host arrays): upload, one MC step, result back
    r_now, m_now = arm.download()
    box_now = np.array(arm.observables()["box"], dtype=np.float64) if not sharded else box  # (the sharded runs are NVT)
    r_pin = torch.from_numpy(np.ascontiguousarray(r_now)).pin_memory()
    m_pin = torch.from_numpy(np.ascontiguousarray(m_now.reshape(-1, 9))).pin_memory()
    r_host, m_host = r_pin.numpy(), m_pin.numpy()
    def e2e_step():
        arm.upload(box_now, r_host, m_host)   # H2D of the step's input configuration + cell build
        if box_moves:
            arm.total_bonds()                 # the initial energy a box move is weighed against (main.c:55-61)
        arm.sweep(1)                          # one MC step

    for i in range(2):
        e2e_step()
        arm.observables()
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step()
        arm.observables()                     # D2H of the step's result (energy, counters)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(getattr(arm.drv, "h2d_bytes", 0)) or int(r_host.nbytes + m_host.nbytes)
    d2h = 96
    # round trip: the whole state comes back to the host every step (pp_download)
    n_rt = max(2, n_e2e // 2)
    back = None
    if not sharded:  # the state lands in pinned host arrays (a slab driver gathers through its own buffers)
        r_back = torch.empty((max(N, 1), 3), dtype=torch.float64).pin_memory()
        m_back = torch.empty((max(N, 1), 9), dtype=torch.float64).pin_memory()
        back = (r_back.numpy(), m_back.numpy())
        arm.download(back)
    barrier()
    t0 = time.perf_counter()
    for i in range(n_rt):
        e2e_step()
        arm.download(back)
    barrier()
    rt_s = time.perf_counter() - t0

    # ---- max over ranks ------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([total_ms, warm_ms, e2e_s, rt_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, warm_ms, e2e_s, rt_s = [float(x) for x in t.tolist()]
        if sharded:
            hsum = torch.tensor([float(h2d)], device=dev, dtype=torch.float64)
            dist.all_reduce(hsum, op=dist.ReduceOp.SUM)
            h2d = int(hsum.item())
    value = K * attempts_per_step / (total_ms * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and cpu_budget_s > 0:
        res = time_reference(w, reference_sample_sweeps(w, cpu_budget_s), 0, local_rank)
        if res is not None:
            cpu = {"value": res["value"], "unit": "attempts/s", "cores": 1, "kind": "reference",
                   "sample": f"{res['sweeps']} MC steps of the unmodified reference ({res['flags']}) on the same configuration file, "
                             "wall(steps) - wall(0 steps); single-threaded code"}
    arm.close()
    if rank != 0:
        return None
    if sharded:
        par = (f"one system in {world} x-slabs; accepted moves of the boundary layers are stored into the neighbour GPU's halo copy over NVLink "
               "from inside the sweep kernel, phase flags in peer memory; migration every shift period through peer mailboxes")
    else:
        par = "single GPU" if world == 1 else f"{world} independent replicas"
    line = {
        "metric": "MC roto-translation attempts/sec", "value": value, "unit": "attempts/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if sharded or world == 1 else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(w), "N_total": int(N * n_systems), "parallelism": par,
                   "configuration": data_desc, "l2": "flushed between steps (256 MiB memset), per-step CUDA events on the launch stream",
                   "lanes_per_cell": eng.params.lanes_per_cell or "auto", "cells": cells},
        "value_l2_warm": k_warm * attempts_per_step / (warm_ms * 1e-3),
        "wall_s_timed_region": t_wall,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": SWEEP_TRAFFIC.get((w, world)),
                     "kernel": sweep_kernel, "peak_source": peak_src + " (of measured)",
                     "note": "algorithmic bytes = 64 B per attempted move (SURVEY 8d); traffic = ncu dram bytes per launch (profiles/); "
                             "the sweep is latency / issue bound, not HBM bound"},
        "kernels": kernels,
        "pair_checks_per_s_l2_warm": pair_checks_per_s,
        "cpu_baseline": cpu,
        "e2e": {"value": n_e2e * attempts_per_step / e2e_s, "unit": "attempts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "step": "upload (pinned host configuration -> HBM) + one MC step (cell build on the step's shifted grid, proposals, eight colour passes) + observables read-back", "steps": n_e2e},
        "e2e_roundtrip": {"value": n_rt * attempts_per_step / rt_s, "unit": "attempts/s", "h2d_bytes_per_step": h2d,
                          "d2h_bytes_per_step": int(r_host.nbytes + m_host.nbytes),
                          "step": "upload + one MC step + pp_download of every position and orientation", "steps": n_rt},
        "gpu_launches": int(launches_timed) * (world if sharded else 1),
        "clocks": clocks,
        "energy_check": {"bonds": b_chk, "ok": bool(ok)},
    }
    if box_moves:
        kind = 6 if d["extra"].get("ensemble") == 5 else 7
        tries = o1["tries"][kind] - o0["tries"][kind]
        acc = o1["accepted"][kind] - o0["accepted"][kind]
        line["box_moves"] = {"kind": "volume (MC.c:450-513)" if kind == 6 else "Lx (MC.c:515-585)", "tries": int(tries), "accepted": int(acc),
                             "per_s": tries / (total_ms * 1e-3), "steps_per_s": K / (total_ms * 1e-3),
                             "note": "each box move = K4 rescale + K3 full bond count + host Metropolis (one host round trip)"}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas of the workload instead of one sharded system")
    ap.add_argument("--lanes", type=int, default=0, help="tile width of the tile kernel (0: auto, 1: thread-per-cell kernel wherever it fits)")
    ap.add_argument("--cell-lanes", type=int, default=0, help="lanes per cell of the thread-per-cell kernel (0: auto, 1, 2, 4, 8)")
    ap.add_argument("--cells-max-occupancy", type=int, default=-1,
                    help="cells with at most this many particles go to the thread-per-cell kernel (-1: auto, 0: tile kernel only)")
    ap.add_argument("--cells-mid-occupancy", type=int, default=-1,
                    help="cells with more particles than --cells-max-occupancy and at most this many run on 4 lanes per cell (-1: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="one GPU: only the headline workload, not the other BASELINE configurations")
    ap.add_argument("--e2e-steps", type=int, default=None)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    w = args.workload if args.workload != "auto" else DEFAULT_WORKLOAD
    big = WORKLOADS[w]["N"] > 500000
    if args.steps is None:
        args.steps = 200 if big else 2000
    if args.warmup is None:
        args.warmup = 10 if big else 50
    if args.e2e_steps is None:
        args.e2e_steps = 5 if big else 100
    if args.impl == "reference":
        return run_reference_arm(args, w)

    import torch
    import patchyparticles_b200 as pkg

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = (rank, local_rank, world, dev, barrier, dist)
    K, W = args.steps, max(args.warmup, 3)
    line = measure(pkg, args, w, K, W, max(3, min(args.e2e_steps, K)), 0. if args.no_cpu_baseline else 15., ctx, True)
    if world == 1 and not args.no_configs and args.workload == "auto":
        extra = {}
        for name in EXTRA_CONFIGS:
            n = WORKLOADS[name]["N"]
            k = 200 if n > 100000 else (400 if n > 10000 else 1000)
            try:
                sub = measure(pkg, args, name, k, 20, 20, 0. if args.no_cpu_baseline else 6., ctx, False)
            except Exception as exc:  # a secondary configuration must not take the headline down
                extra[name] = {"error": repr(exc)}
                continue
            keep = ("value", "ms_per_step", "steps", "warmup", "value_l2_warm", "roofline", "kernels", "cpu_baseline", "e2e", "e2e_roundtrip",
                    "gpu_launches", "box_moves", "energy_check", "pair_checks_per_s_l2_warm")
            extra[name] = {"workload": workload_name(name), "configuration": sub["config"]["configuration"], "cells": sub["config"]["cells"]}
            extra[name].update({k_: sub[k_] for k_ in keep if k_ in sub})
        line["configs"] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
