#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the FDTD time-stepping hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json config #5 = weak-scaling sweep, N = 1 is config #3's grid): phononic
crystal, square lattice of full-depth Au cylinders (pitch 32, r = 8) in GaAs, (512*N) x 512 x 512
points, absorbing + free-surface boundaries, left-wall sin source f = 100, courant 0.1, fp64
(the reference's dtype), FAST arithmetic, x-slab of 512 planes per GPU.  A "step" is one time
step of the whole grid.  Synthetic, deterministic; all arrays are far larger than L2 (126 MB).

One JSON line on rank 0:
  value      Gcell-updates/s, device-timed (CUDA events on the launching stream, max over ranks),
             fields resident in HBM
  e2e        the same metric through the plugin API (Solver.init + Solver.run, reference interface),
             wall clock of run(): per step the source sample goes host->device and the recorded
             surface plane (uz at z-index 0, BASELINE config #3; --e2e-fields ux,uy,uz for all three)
             comes device->pinned host->HDF5 file
  roofline   dominant kernel (k_step_march): algorithmic bytes (73 B/cell fp64: 9 field words + 1
             class byte) / measured kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the NumPy oracle (restatement of the reference's NumPy solver) on a bounded sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC, UNIT = "Gcell-updates/s", "Gcell/s"
NY = NZ = 512
NX_PER_GPU = 512
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback"


def b_alg(dtype):
    return 9 * (8 if dtype == "f64" else 4) + 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (nvidia-smi -lms, the
    recipe's clocks line), started before and stopped after it."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)      # let the first samples arrive
        except Exception:
            self.proc = None

    def finish(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.05)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[c.strip() for c in ln.split(",")] for ln in out.splitlines() if ln.strip()]
        num = lambda v: v.replace(".", "", 1).isdigit()
        sm = [float(r[0]) for r in rows if r and num(r[0])]
        mx = [float(r[1]) for r in rows if len(r) > 1 and num(r[1])]
        pw = [float(r[2]) for r in rows if len(r) > 2 and num(r[2])]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        # samples under load = the upper half by power draw (the sampler brackets the timed region)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


# ---------------------------------------------------------------------------------------------
# CPU legs (oracle = checker; only used here as the timed CPU baseline, never on the product path)
# ---------------------------------------------------------------------------------------------
def cpu_sample(threads, steps, warmup, n=128):
    """The reference's NumPy algorithm (oracle/fdtd_numpy.py) on an n^3 block of the same crystal
    (same lattice / materials / source / boundaries).  Returns (gcells_per_s, description)."""
    from oracle import fdtd_numpy as onp
    from phonomena_b200.workloads import crystal_case
    c = crystal_case(n, n, n)
    C, P = onp.set_constants(c.x, c.y, c.z, onp.make_targets(c.targets.tolist()), c.prim_c, c.prim_p, c.sec_c, c.sec_p)
    o = onp.OracleSolver(c.x, c.y, c.z, C, P, c.dt, wave="sin", wave_args={"f": 100}, threads=threads)
    for _ in range(warmup):
        o.step()
    t0 = time.perf_counter()
    for _ in range(steps):
        o.step()
    dt = time.perf_counter() - t0
    o.close()
    return n ** 3 * steps / dt / 1e9, "%d^3 block of the same crystal, %d steps, NumPy float64, %d thread(s)" % (n, steps, threads)


def cpu_sample_c(steps, warmup, n=256):
    """Context only: the C restatement (oracle/fdtd_c.c, bit-identical to the NumPy one) with OpenMP on every
    host core -- a far stronger CPU program than the reference's NumPy, reported beside it, never as the baseline."""
    from oracle import fdtd_c, fdtd_numpy as onp
    from phonomena_b200.workloads import crystal_case
    c = crystal_case(n, n, n)
    ids = onp.material_id_map(c.x, c.y, c.z, onp.make_targets(c.targets.tolist()))
    o = fdtd_c.COracle(c.x, c.y, c.z, ids, [c.prim_c, c.sec_c], [c.prim_p, c.sec_p], c.dt, wave="sin", wave_args={"f": 100}, omp=True)
    o.run(warmup)
    t0 = time.perf_counter()
    o.run(steps)
    dt = time.perf_counter() - t0
    o.close()
    cores = int(os.environ.get("OMP_NUM_THREADS", 0)) or len(os.sched_getaffinity(0))
    return {"value": n ** 3 * steps / dt / 1e9, "unit": UNIT, "cores": cores,
            "sample": "%d^3 block of the same crystal, %d steps, C + OpenMP (gcc -O2 -ffp-contract=off), float64" % (n, steps)}


def workload_name(n):
    """The workload both arms report (the reference arm times a bounded sample of it)."""
    return ("phononic crystal %dx%dx%d (Au cylinders pitch 32 r 8 in GaAs; BASELINE config #%s), x-slabs of %d planes/GPU, "
            "Mur ABC + free surface + left-wall sin source f=100, courant 0.1" % (NX_PER_GPU * n, NY, NZ, "3" if n == 1 else "5", NX_PER_GPU))


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path = NumPy slicing
    arithmetic, timed through the oracle port (the reference is pure Python and does not travel to
    the GPU box; the port is pinned bit-for-bit to it, tests/test_oracle_golden.py).  All host
    threads the algorithm can use: the six stress / three displacement tasks of solver_threading."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = max(1, min(6, len(os.sched_getaffinity(0))))
    steps, warmup = max(1, min(args.steps, 200)), max(1, min(args.warmup, 10))    # ~0.12 s per 128^3 step: K = 50 is 6 s
    v, sample = cpu_sample(threads, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 128 ** 3 / v / 1e6, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(max(1, args.gpus)), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)), "numpy": np.__version__},
    }
    try:
        line["extra"] = {"c_openmp_port": cpu_sample_c(min(steps, 10), 1)}
    except Exception as exc:      # context only
        line["extra"] = {"c_openmp_port": {"error": str(exc)[:200]}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    from phonomena_b200 import _lib, hostmath as hm
    from phonomena_b200.workloads import crystal_case

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    n = max(world, 1)
    if args.gpus != n and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_
        torch.cuda.set_device(local)
        dist_.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_

    nx = NX_PER_GPU * n
    case = crystal_case(nx, NY, NZ)
    x0, nxl = hm.split_slabs(nx, n)[rank]
    K, W = args.steps, args.warmup
    dtype, arith = args.dtype, args.arith
    e = case.make_engine(steps=W + K, x0=x0, nxl=nxl, dtype=dtype, arith=arith, device=local, kernel=args.kernel)
    halo = "none"
    if world > 1:
        def allgather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        def broadcast(obj):
            box = [obj]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        halo = e.connect(rank, world, allgather, broadcast)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    e.run(W)
    e.sync()
    e.profile(1)
    l0 = e.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    ms = e.run_timed(K)
    barrier()
    clocks = sampler.finish() if sampler else None
    launches = e.launch_count - l0
    kms, kn = e.profile(0)
    if dist is not None:
        import torch
        t = torch.tensor([ms, kms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kms = float(t[0]), float(t[1])
        tl = torch.tensor([launches], dtype=torch.int64, device="cuda")
        dist.all_reduce(tl)
        launches = int(tl[0])
    info = e.info()
    cells = nx * NY * NZ
    value = cells * K / ms / 1e6            # Gcell/s, whole job
    e.close()

    # ---- e2e through the plugin API (reference interface), per rank its slab ------------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(case, n, rank, local, K, dtype, arith, dist, fields=tuple(args.e2e_fields.split(",")))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = hbm_peak()
    cells_local = nxl * NY * NZ
    # the stencil launches of the K timed steps process cells_local * K cell updates in total (with slabs a step
    # is an edge launch per neighbour + one interior launch); kms is their summed device time on the slowest rank
    ach = b_alg(dtype) * cells_local * K / (kms * 1e-3) / 1e9 if kms > 0 else None     # GB/s of one GPU's kernel
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("%s_512" % dtype, {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload_name(n),
                   "arith": arith, "material": "indexed (1-byte stencil class)", "kernel": info["kernel"],
                   "l2": "inputs (%.1f GB/GPU) far exceed the 126 MB L2; no explicit flush" % (info["device_bytes"] / 1e9),
                   "halo": {"none": "none", "nccl": "NCCL send/recv of 3 planes per direction per step, overlapped with the interior update",
                            "p2p": "fused: the stencil kernel stores its edge planes into the neighbours' ghost planes over NVLink (CUDA IPC), "
                                   "stream-ordered flag write/wait, no collective call"}[halo]},
        "clocks": clocks, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "k_step_march", "kernel_ms_per_step": kms / K if K else None, "kernel_launches_per_step": kn / K if K else None,
                     "algorithmic_bytes_per_cell": b_alg(dtype), "kernel_share_of_step": kms / ms if ms else None},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if n == 1 and not args.no_cpu:
        v, sample = cpu_sample(1, 5, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample}
    if args.extra:
        line["extra"] = extra_runs(case, local)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_e2e(case, n, rank, local, K, dtype, arith, dist, fields=("ux", "uy", "uz")):
    """Solver.init + Solver.run through the plugin with host buffers: mesh lines, inclusion list
    and tables go host->device in init(); inside the timed run() every chunk's source samples go
    host->device and every step's surface planes come back through the pinned ring into the HDF5
    file.  Timed region = run() (the reference times its loop the same way, base_solver.py:239-279)."""
    import tempfile
    from phonomena_b200.solver_b200 import Solver
    g, m = case.as_grid_material()
    s = Solver()
    s.cfg.update({"precision": {"f64": "fp64", "f32": "fp32"}[dtype], "arith": arith, "device": local, "wave": "sin",
                  "wave_args": {"f": 100}, "write_mode": "thread", "record": "surface", "record_every": 1,
                  "slabs_from_env": n > 1, "chunk_steps": 10,
                  "record_fields": list(fields),
                  "merge_slabs": False})      # N > 1: one slab file per rank (concatenating them is post-processing)
    # output file: tmpfs when the box has one with room (the run measures the solver and its copies, not the
    # scratch disk of the box), else the temp directory; reported in e2e["file_dir"]
    out_dir = tempfile.gettempdir()
    try:
        st = os.statvfs("/dev/shm")
        if st.f_bavail * st.f_frsize > 4 * (1 << 30) * max(1, n):
            out_dir = "/dev/shm"
    except OSError:
        pass
    s.file = os.path.join(out_dir, "phb_bench_rank%d.h5" % rank)
    steps = max(K, 10)
    s.init(g, m, steps)
    # warm-up (untimed), as for `value`: init() is host-bound for seconds (mesh / density to the file), the GPU drops to
    # idle clocks meanwhile -- step a small separate engine on the same device for >= 0.2 s before the timed run()
    from phonomena_b200.workloads import crystal_case
    with crystal_case(192, 192, 192).make_engine(steps=20000, dtype=dtype, arith=arith, device=local) as warm:
        t_w, n_w = time.perf_counter(), 0
        while time.perf_counter() - t_w < 0.2 and n_w + 50 <= 20000:
            warm.run(50)
            warm.sync()
            n_w += 50
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    s.run()
    dt = time.perf_counter() - t0
    if dist is not None:
        import torch
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    nx, ny, nz = case.shape
    path = s.file if n == 1 else "%s.rank%d" % (s.file, rank)
    out = {"value": nx * ny * nz * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8,
           "d2h_bytes_per_step": 8 * sum(v for k, v in (("ux", (nx - 1) * ny), ("uy", nx * (ny - 1)), ("uz", nx * ny)) if k in fields),
           "steps": steps,
           "what": "Solver.run(): source sample H2D + surface %s planes D2H (pinned ring) -> HDF5 every step" % ",".join(fields)
                   + ("; one slab file per rank, max over ranks" if n > 1 else ""),
           "file_bytes": os.path.getsize(path), "file_dir": out_dir,
           "writer_finish_ms": 1e3 * s.stats.get("writer_finish_seconds", 0.0),
           "writer_write_ms": 1e3 * s.stats.get("writer_write_seconds", 0.0),
           "writer_wait_ms": 1e3 * s.stats.get("writer_wait_seconds", 0.0), "run_ms": 1e3 * dt,
           "loop_ms": 1e3 * s.stats.get("loop_seconds", 0.0)}
    try:
        os.remove(path)
    except OSError:
        pass
    s._close_engine()
    return out


def extra_runs(case, local):
    """fp32 and EXACT-arithmetic numbers on the same grid (reported beside the headline)."""
    out = {}
    for dtype, arith in (("f32", "fast"), ("f64", "exact")):
        e = case.make_engine(steps=13, dtype=dtype, arith=arith, device=local)
        e.run(3)
        e.sync()
        ms = e.run_timed(10)
        nx, ny, nz = case.shape
        out["%s_%s" % (dtype, arith)] = {"gcells_per_s": nx * ny * nz * 10 / ms / 1e6, "ms_per_step": ms / 10,
                                         "frac_of_hbm_roofline": nx * ny * nz * 10 / ms / 1e6 * b_alg(dtype) / hbm_peak()[0]}
        e.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--arith", default="fast", choices=["fast", "exact"])
    ap.add_argument("--kernel", default="auto")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-fields", default="uz",
                    help="surface components the e2e run records every step: BASELINE config #3 is 'surface u_z HDF5 recording' "
                         "(default); ux,uy,uz = everything the reference's Writer stores at the surface")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--extra", action="store_true", help="also time fp32 and fp64-exact on the same grid")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
