#!/usr/bin/env python
"""bench.py -- attempted roto-translation moves per second of the checkerboard MC engine.

    python bench.py --gpus N --steps K --warmup W            (our arm)
    python bench.py --impl reference --gpus N --steps K --warmup W   (reference CPU arm)

A "step" is one MC step of the reference (main.c:119): N_particles attempted roto-translations
(one 8-colour checkerboard sweep including its share of the cell builds).  The workload is the same
at every GPU count -- BASELINE.json configs[3], N = 4 194 304 at rho = 0.6, NVT -- on one engine at
--gpus 1 and split into N x-slabs at --gpus N (strong scaling), so the driver's efficiency divides
like by like.  The one-GPU line also carries, under "configs", the other BASELINE configurations
measured in the same run: N = 65536 (configs[1]), the NPT run at N = 32768 (configs[2]), the FCC
crystal at N = 256 000 (configs[4]) and N = 1000 (configs[0]).  --workload picks another one as the headline.

Prints ONE JSON line on rank 0 (see the task contract): value = device-timed throughput with the
state resident in HBM (L2 flushed between steps), e2e = the same metric through the C ABI with
host buffers (H2D of the configuration + sweep + D2H of the observables every step), roofline for
the sweep kernel from live CUDA-event timings, cpu_baseline = the unmodified reference timed on
this box's host cores.
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

WORKLOADS = {
    # name: (N, rho, temperature)
    "64k": (65536, 0.5, 0.1),
    "1k": (1000, 0.1, 0.1),
    "256k": (262144, 0.5, 0.1),
    "512k": (524288, 0.6, 0.1),
    "1M": (1048576, 0.6, 0.1),
    "4M": (4194304, 0.6, 0.1),
}
DEFAULT_WORKLOAD = "4M"
# dram__bytes_read.sum + dram__bytes_write.sum per colour-pass launch from the committed ncu captures (profiles/), by
# (workload, GPUs); None where no capture of that shape is committed
SWEEP_TRAFFIC = {}
KF_DELTA, KF_COSMAX, DISP_MAX, THETA_MAX, SEED = 0.119, 0.98, 0.1, 0.1, 12345
SWEEP_BYTES_PER_ATTEMPT = 64   # state read once + written once per sweep at S = 32 B
CELLS_BYTES_PER_PARTICLE = 72
BONDS_BYTES_PER_PARTICLE = 32
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def workload_name(w):
    N, rho, T = WORKLOADS[w]
    return f"N={N} rho={rho} NVT T={T} KF_delta={KF_DELTA} KF_cosmax={KF_COSMAX} Disp_max=Theta_max={DISP_MAX}"


# ------------------------------------------------------------------------------------------------
# configuration: the reference's own deterministic generator when it is available, else a seeded fluid
# ------------------------------------------------------------------------------------------------
def reference_rrr(N, rho):
    """Path of generated.rrr from the reference's generator (generator.c never seeds drand48)."""
    gen = os.path.join(REF_DIR, "generator")
    if not os.path.exists(gen) or N > 300000:
        return None
    cache = os.path.join(tempfile.gettempdir(), f"pp_generated_{N}_{rho}.rrr")
    if not os.path.exists(cache):
        d = tempfile.mkdtemp(prefix="ppgen_")
        try:
            subprocess.run([gen, str(N), str(rho)], cwd=d, check=True, stderr=subprocess.DEVNULL)
            shutil.move(os.path.join(d, "generated.rrr"), cache)
        except Exception:
            return None
        finally:
            shutil.rmtree(d, ignore_errors=True)
    return cache


def make_configuration(w):
    import patchyparticles_b200 as pkg
    N, rho, _ = WORKLOADS[w]
    path = reference_rrr(N, rho)
    if path is not None:
        _, box, r, m = pkg.read_rrr(path)
        return box, r, m, f"reference generator {N} {rho}", path
    box, r, m = pkg.random_fluid(N, rho, seed=1)
    return box, r, m, f"seeded lattice-thinned hard-sphere fluid {N} {rho}", None


# ------------------------------------------------------------------------------------------------
# reference CPU arm
# ------------------------------------------------------------------------------------------------
def _run_ref(binary, workdir, rrr, steps, T):
    inp = os.path.join(workdir, "input")
    with open(inp, "w") as f:
        f.write(f"Dynamics = 0\nEnsemble = 0\nTemperature = {T}\nSteps = {steps}\nDisp_max = {DISP_MAX}\n"
                f"Theta_max = {THETA_MAX}\nInitial_conditions_file = {rrr}\nPrint_every = 1000000000\n"
                f"Save_every = 1000000000\nConfiguration_folder = .\nKF_delta = {KF_DELTA}\nKF_cosmax = {KF_COSMAX}\n"
                f"Seed = {SEED}\nLog_file = none\n")
    t0 = time.perf_counter()
    p = subprocess.run([binary, "input"], cwd=workdir, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return time.perf_counter() - t0, p.returncode


def time_reference(w, sweeps, warmup=0):
    """wall(sweeps) - wall(0 sweeps) of the unmodified reference binary, single thread (the reference
    has no threads: main.c:26-143); `warmup` steps are run once before, untimed (page cache, CPU clocks)."""
    N, rho, T = WORKLOADS[w]
    rrr = reference_rrr(N, rho)
    if rrr is None:
        # no generator output (large N): write our synthetic configuration as .rrr
        import patchyparticles_b200 as pkg
        import _oracle  # test-only C writer (output.c:302-317 format); allowed in the reference arm
        box, r, m = pkg.random_fluid(N, rho, seed=1)
        rrr = os.path.join(tempfile.gettempdir(), f"pp_synth_{N}_{rho}.rrr")
        if not os.path.exists(rrr):
            _oracle.write_rrr(rrr, 0, box, r, m.reshape(-1, 9))
    d = tempfile.mkdtemp(prefix="ppref_")
    try:
        kind_flags = None
        for name, flags in (("PatchyParticles_native", "-O3 -march=native (built on the build host)"),
                            ("PatchyParticles", "-O3 -march=x86-64-v3")):
            binary = os.path.join(REF_DIR, name)
            if not os.path.exists(binary):
                continue
            t_init, rc = _run_ref(binary, d, rrr, 0, T)
            if rc != 0:
                continue
            kind_flags = flags
            break
        if kind_flags is None:
            return None
        if warmup > 0:
            _run_ref(binary, d, rrr, warmup, T)
        t_run, rc = _run_ref(binary, d, rrr, sweeps, T)
        if rc != 0:
            return None
        dt = max(t_run - t_init, 1e-9)
        return {"value": sweeps * N / dt, "seconds": dt, "init_seconds": t_init, "flags": kind_flags, "sweeps": sweeps}
    finally:
        shutil.rmtree(d, ignore_errors=True)


def reference_sample_sweeps(w, budget_s=15.0):
    N = WORKLOADS[w][0]
    est = N / (4.0e5 if N < 500000 else 1.5e5)  # attempts/s/core measured at survey time (cache-bound at 4M)
    return max(1, int(budget_s / est))


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N = WORKLOADS[w][0]
    sweeps = max(1, min(args.steps, reference_sample_sweeps(w, 40.0)))
    warm = max(0, min(args.warmup, 3, sweeps))  # a separate, untimed run of the binary
    res = time_reference(w, sweeps, warm)
    if res is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref reference binaries missing or not runnable on this host"}))
        return 0
    line = {
        "impl": "reference", "metric": "MC roto-translation attempts/sec", "value": res["value"], "unit": "attempts/s",
        "n_gpus": args.gpus, "steps": sweeps, "warmup": warm, "ms_per_step": 1e3 * res["seconds"] / sweeps,
        "higher_is_better": True, "scaling": "strong" if WORKLOADS[w][0] > 500000 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": workload_name(w), "N_total": N},
        "cpu_baseline": {"value": res["value"], "unit": "attempts/s", "cores": 1, "kind": "reference",
                         "sample": f"{sweeps} MC steps of the unmodified reference ({res['flags']}), wall(steps) - wall(0 steps); "
                                   "the reference is single-threaded"},
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
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
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


class Arm:
    """The engine under test: one Engine (world == 1 or independent replicas) or one SlabEngine rank."""

    def __init__(self, pkg, args, w, rank, local_rank, world, sharded):
        import torch
        N, rho, T = WORKLOADS[w]
        params = dict(temperature=T, kf_delta=KF_DELTA, kf_cosmax=KF_COSMAX, disp_max=DISP_MAX, theta_max=THETA_MAX,
                      lanes_per_cell=args.lanes, cells_max_occupancy=args.cells_max_occupancy, cell_lanes=args.cell_lanes,
                      cells_mid_occupancy=args.cells_mid_occupancy)
        self.sharded = sharded
        if sharded:
            from patchyparticles_b200.slab import SlabEngine
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

    def pair_checks(self):
        o = self.drv.observables()
        return o["pair_checks"] if self.sharded else o.pair_checks

    def result_readback(self):
        """D2H of the step's result: the running energy and the counters."""
        return self.drv.observables()

    def check_energy(self):
        return self.drv.check_energy()

    def download(self):
        return self.drv.download()

    def cells(self):
        return list(self.eng.observables().n_side)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--replicas", action="store_true", help="N > 1: independent replicas of the 1-GPU workload instead of one sharded system")
    ap.add_argument("--lanes", type=int, default=0, help="tile width of the tile kernel (0: auto, 1: thread-per-cell kernel wherever it fits)")
    ap.add_argument("--cell-lanes", type=int, default=0, help="lanes per cell of the thread-per-cell kernel (0: auto, 1, 2, 4, 8)")
    ap.add_argument("--cells-max-occupancy", type=int, default=-1,
                    help="cells with at most this many particles go to the thread-per-cell kernel (-1: auto, 0: tile kernel only)")
    ap.add_argument("--cells-mid-occupancy", type=int, default=-1,
                    help="cells with more particles than --cells-max-occupancy and at most this many run on 4 lanes per cell (-1: auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--relax-sweeps", type=int, default=300,
                    help="untimed sweeps that turn the lattice-thinned synthetic configuration into a fluid before anything is measured")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # N = 1: BASELINE.json configs[1] (N = 65536).  N > 1: configs[3], one N = 4M system sharded into x slabs (strong scaling)
    sharded = world > 1 and not args.replicas
    w = args.workload if args.workload != "auto" else ("4M" if sharded else "64k")
    big = WORKLOADS[w][0] > 500000
    if args.steps is None:
        args.steps = 200 if big else 2000
    if args.warmup is None:
        args.warmup = 10 if big else 50
    if args.e2e_steps is None:
        args.e2e_steps = 5 if big else 200
    if args.impl == "reference":
        return run_reference_arm(args, w)

    import numpy as np
    import torch
    import patchyparticles_b200 as pkg

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    N, rho, T = WORKLOADS[w]
    K, W = args.steps, max(args.warmup, 3)
    box, r, m, data_desc, rrr_path = make_configuration(w)
    stream = torch.cuda.current_stream()
    arm = Arm(pkg, args, w, rank, local_rank, world, sharded)
    eng = arm.eng
    arm.upload(box, r, m)
    bonds0, ov0 = arm.total_bonds()
    assert ov0 == 0, "initial configuration overlaps"
    n_systems = 1 if sharded else world          # independent systems in flight
    attempts_per_step = float(N) * n_systems      # whole-job attempts per step

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    # ---- untimed relaxation of the lattice-thinned synthetic configuration into a fluid ------------------
    if rrr_path is None and args.relax_sweeps > 0:
        arm.sweep(args.relax_sweeps, sync=True)
        r, m = arm.download()
        data_desc += f", relaxed by {args.relax_sweeps} untimed sweeps"

    # ---- warm-up ----------------------------------------------------------------------------
    arm.sweep(W, sync=True)
    flush.zero_()
    barrier()

    # ---- timed: K steps, L2 flushed between steps, per-step CUDA events on the launch stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    eng.reset_kernel_stats()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
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

    # ---- same K steps back to back without the flush (steady state of a real run) -----------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    checks0 = arm.pair_checks()
    barrier()
    ev0.record(stream)
    arm.sweep(K)
    ev1.record(stream)
    barrier()
    warm_ms = ev0.elapsed_time(ev1)
    pair_checks_per_s = (arm.pair_checks() - checks0) * (1 if sharded else world) / (warm_ms * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel timing for the roofline (CUDA events around each kernel class) -----------------
    eng.set_profiling(True)
    eng.reset_kernel_stats()
    n_prof = min(K, 300)
    for i in range(n_prof):
        flush.zero_()
        arm.sweep(1)
    arm.total_bonds()
    eng.sync()
    stats = eng.kernel_stats()
    eng.set_profiling(False)
    n_local = eng.slab_range()[3] if sharded else N
    peak, peak_src = measured_peaks()
    sweep_launches, sweep_ms = stats["sweep"]
    n_passes = 8 * n_prof
    # hybrid routing: two concurrent launches per colour pass (sparse cells / crowded cells), timed together
    launches_per_pass = max(1, round(sweep_launches / n_passes))
    per_launch_ms = sweep_ms / n_passes
    sweep_kernel = "k_sweep_pass (tile kernel)" if n_local < 150000 and args.cells_max_occupancy < 0 and args.lanes != 1 else "k_sweep_fused (8 / 4 / 1 lanes per cell in one launch)"
    n_owned = N / world if sharded else N
    bytes_per_launch = SWEEP_BYTES_PER_ATTEMPT * n_owned / 8.0
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    cell_builds = max(stats["cells"][0] // 10, 1)
    kernels = {
        "sweep_pass": {"launches": sweep_launches, "passes": n_passes, "ms_per_launch": per_launch_ms, "alg_bytes_per_launch": bytes_per_launch,
                       "GBps": achieved},
        "cell_build": {"launches": stats["cells"][0], "builds": cell_builds, "ms_per_build": stats["cells"][1] / cell_builds,
                       "GBps": CELLS_BYTES_PER_PARTICLE * n_local / (stats["cells"][1] / cell_builds * 1e-3) / 1e9},
        "proposals": {"launches": stats["rng"][0], "ms_per_launch": stats["rng"][1] / max(stats["rng"][0], 1)},
        "bond_count": {"launches": stats["bonds"][0], "ms": stats["bonds"][1],
                       "GBps": BONDS_BYTES_PER_PARTICLE * n_owned / (max(stats["bonds"][1], 1e-9) * 1e-3) / 1e9},
    }
    if sharded:  # pushes into the neighbour's mailbox + pulls (the pull includes waiting for the neighbour)
        kernels["halo"] = {"launches": stats["halo"][0], "ms_per_sweep": stats["halo"][1] / max(n_prof, 1)}

    # invariant after the measured regions: running energy == recount (MC_check_energy)
    b_chk, ov_chk, ok = arm.check_energy()
    assert ok, "energy check failed after the timed region"

    # ---- e2e through the public API with host buffers ---------------------------------------------
    r_pin = torch.from_numpy(np.ascontiguousarray(r)).pin_memory()
    m_pin = torch.from_numpy(np.ascontiguousarray(m.reshape(-1, 9))).pin_memory()
    r_host, m_host = r_pin.numpy(), m_pin.numpy()
    n_e2e = max(3, min(args.e2e_steps, K))
    for i in range(2):
        arm.upload(box, r_host, m_host)
        arm.sweep(1)
        arm.result_readback()
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        arm.upload(box, r_host, m_host)       # H2D of the step's input configuration + cell build
        arm.sweep(1)                          # one MC step
        arm.result_readback()                 # D2H of the step's result (energy, counters)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = int(r_host.nbytes + m_host.nbytes)
    d2h = 96

    # ---- max over ranks ------------------------------------------------------------------------
    if world > 1:
        t = torch.tensor([total_ms, warm_ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, warm_ms, e2e_s = [float(x) for x in t.tolist()]
    value = K * attempts_per_step / (total_ms * 1e-3)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = time_reference(w, reference_sample_sweeps(w, 15.0))
        if res is not None:
            cpu = {"value": res["value"], "unit": "attempts/s", "cores": 1, "kind": "reference",
                   "sample": f"{res['sweeps']} MC steps of the unmodified reference ({res['flags']}) on the same configuration, "
                             "wall(steps) - wall(0 steps); single-threaded code"}

    if rank == 0:
        if sharded:
            par = (f"one system in {world} x-slabs; halo refresh (2 per sweep) and migration (every shift period) as stores into the "
                   "neighbour GPU's memory over NVLink (CUDA IPC mailboxes, device-side flags)")
        else:
            par = "single GPU" if world == 1 else f"{world} independent replicas"
        line = {
            "metric": "MC roto-translation attempts/sec", "value": value, "unit": "attempts/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(w), "N_total": int(N * n_systems), "parallelism": par,
                       "configuration": data_desc, "l2": "flushed between steps (256 MiB memset), per-step CUDA events on the launch stream",
                       "lanes_per_cell": eng.params.lanes_per_cell or "auto", "cells": arm.cells()},
            "value_l2_warm": K * attempts_per_step / (warm_ms * 1e-3),
            "wall_s_timed_region": t_wall,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": SWEEP_TRAFFIC.get((w, world)),
                         "kernel": sweep_kernel, "peak_source": peak_src + " (of measured)",
                         "note": "algorithmic bytes = 64 B per attempted move (SURVEY 8d); traffic = ncu dram bytes per launch (profiles/); "
                                 "the sweep is latency / L2 bound, not HBM bound"},
            "kernels": kernels,
            "pair_checks_per_s_l2_warm": pair_checks_per_s,
            "cpu_baseline": cpu,
            "e2e": {"value": n_e2e * attempts_per_step / e2e_s, "unit": "attempts/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "step": "pp_upload (pinned host configuration -> HBM, cell build) + one MC step + observables read-back", "steps": n_e2e},
            "gpu_launches": int(launches_timed) * (world if sharded else 1),
            "clocks": clocks,
            "energy_check": {"bonds": b_chk, "ok": bool(ok)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    eng.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
