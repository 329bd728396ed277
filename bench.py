#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the FDTD time-stepping hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json config #5 = weak-scaling sweep, N = 1 is config #3's grid): phononic
crystal, square lattice of full-depth Au cylinders (pitch 32, r = 8) in GaAs, (512*N) x 512 x 512
points, absorbing + free-surface boundaries, left-wall sin source f = 100, courant 0.1, fp64
(the reference's dtype), FAST arithmetic, x-slab of 512 planes per GPU.  A "step" is one time
step of the whole grid.  Synthetic, deterministic; all arrays are far larger than L2 (126 MB).

--scaling strong: BASELINE.json config #4 instead -- a FIXED 1024 x 512 x 512 crystal split into N x-slabs.

One JSON line on rank 0:
  value      Gcell-updates/s, device-timed (CUDA events on the launching stream, max over ranks),
             fields resident in HBM, production launch path (CUDA-graph replay of the step, slabs included); the kernel
             share for the roofline comes from a SECOND pass of K steps with per-launch event pairs
  parity     N > 1: before the timed region every rank checks its slabs of a small crystal, stepped through
             the same halo exchange, bit for bit against a single-GPU run (phonomena_b200/selfcheck.py);
             a mismatch fails the run (exit code 3)
  e2e        the same metric through the plugin API (Solver.init + Solver.run, reference interface),
             wall clock of run(): per step the source sample goes host->device and the recorded
             surface plane (uz at z-index 0, BASELINE config #3; --e2e-fields ux,uy,uz for all three)
             comes device->pinned host->native writer threads->HDF5 file; over max(K, 200) steps; init() reported as init_s
  roofline   dominant kernel (k_step_march): algorithmic bytes (73 B/cell fp64: 9 field words + 1
             class byte) / measured kernel time, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the UNMODIFIED reference solver (oracle/_ref, staged by `make -C oracle ref`; kind "reference")
             on a bounded sample through its own Solver.init / Solver.run; the NumPy port only if that copy is absent
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
NX_STRONG = 1024              # BASELINE config #4: 1024 x 512 x 512 over 2 / 4 / 8 GPUs
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


def workload_name(n, scaling="weak"):
    """The workload both arms report (the reference arm times a bounded sample of it)."""
    if scaling == "strong":
        return ("phononic crystal %dx%dx%d (Au cylinders pitch 32 r 8 in GaAs; BASELINE config #4), fixed grid in %d x-slab(s) of %d planes, "
                "Mur ABC + free surface + left-wall sin source f=100, courant 0.1" % (NX_STRONG, NY, NZ, n, NX_STRONG // n))
    return ("phononic crystal %dx%dx%d (Au cylinders pitch 32 r 8 in GaAs; BASELINE config #%s), x-slabs of %d planes/GPU, "
            "Mur ABC + free surface + left-wall sin source f=100, courant 0.1" % (NX_PER_GPU * n, NY, NZ, "3" if n == 1 else "5", NX_PER_GPU))


def mem_available_gb():
    try:
        for ln in open("/proc/meminfo"):
            if ln.startswith("MemAvailable"):
                return int(ln.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def reference_sample(steps, n=128, solvers=("solver_threading", "solver_default"), warmup=1):
    """The UNMODIFIED reference (oracle/_ref or /root/reference through oracle/refshim.py) on an n^3 block of the
    bench crystal, through its own Solver.init / Solver.run (oracle/ref_bench.py).  None if no copy is present."""
    from oracle import refshim
    if not refshim.available():
        return None
    from oracle import ref_bench
    return ref_bench.run(n, steps, warmup, solvers), ref_bench.host_info(), refshim.REF_ROOT


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, UNMODIFIED, through its own public
    API (common.importSolver -> Solver.init -> Solver.run, write_mode 'off'; tests/test_speed.py:47-56) on the
    box's host cores.  Headline = solver_threading (6 worker threads, the fastest correct reference solver,
    SURVEY 8d) on a 128^3 block of the bench crystal; solver_default (1 thread) and, memory permitting, both at
    256^3 are reported in `runs`.  Falls back to the NumPy port (pinned bit-for-bit to the reference) only when
    oracle/_ref is absent, and says so (`cpu_baseline.kind`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 20)), 1      # ~0.5 s per 128^3 step of the reference: K = 20 is 10 s
    ref = None
    try:
        ref = reference_sample(steps)
    except Exception as exc:      # a broken copy must not take the arm down: fall back to the port and say why
        sys.stderr.write("reference copy unusable (%s); timing the port instead\n" % exc)
    extra = {}
    if ref is not None:
        runs, host, root = ref
        r = runs["solver_threading"]
        v, threads, kind = r["value"], r["threads"], "reference"
        sample = "128^3 block of the same crystal, %d steps, UNMODIFIED reference solver_threading (6 worker threads) via Solver.init/run, write_mode off" % steps
        extra["runs"] = {"128": runs}
        if mem_available_gb() > 48 and not args.quick_reference:
            try:
                from oracle import ref_bench
                extra["runs"]["256"] = ref_bench.run(256, 3, 0)        # ~17 GB RSS, ~20 s per solver
            except Exception as exc:
                extra["runs"]["256"] = {"error": str(exc)[:200]}
        extra["ref_root"] = "oracle/_ref" if root.endswith("_ref") else root
        try:
            pv, psample = cpu_sample(6, min(steps, 10), 1)
            extra["numpy_port_6_threads"] = {"value": pv, "unit": UNIT, "sample": psample}
        except Exception as exc:
            extra["numpy_port_6_threads"] = {"error": str(exc)[:200]}
    else:
        threads = max(1, min(6, len(os.sched_getaffinity(0))))
        v, sample = cpu_sample(threads, steps, warmup)
        kind = "port"
        from oracle import ref_bench
        host = ref_bench.host_info()
        extra["note"] = "oracle/_ref absent (run `make -C oracle ref` in the build container): NumPy port timed instead"
    try:
        extra["c_openmp_port"] = cpu_sample_c(min(steps, 10), 1)
    except Exception as exc:      # context only
        extra["c_openmp_port"] = {"error": str(exc)[:200]}
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": 128 ** 3 / v / 1e6, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(max(1, args.gpus), args.scaling), "sample": sample},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": host, "extra": extra,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
def run_b200(args):
    from phonomena_b200 import _lib, hostmath as hm, selfcheck
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

    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def broadcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    # ---- N > 1: correctness of the slab path first, through the halo mode the timed run uses -------
    parity = None
    if world > 1 and not args.no_parity:
        parity = selfcheck.slabs_vs_single(rank, world, local, allgather, broadcast)
        if not parity["slabs_bit_identical"]:
            if rank == 0:
                print(json.dumps({"metric": METRIC, "n_gpus": n, "parity": parity, "error": "slabs differ from the single-GPU run"}), flush=True)
            dist.destroy_process_group()
            sys.exit(3)

    strong = args.scaling == "strong"
    nx = NX_STRONG if strong else NX_PER_GPU * n
    case = crystal_case(nx, NY, NZ)
    x0, nxl = hm.split_slabs(nx, n)[rank]
    K, W = args.steps, args.warmup
    dtype, arith = args.dtype, args.arith
    PRIME = 6      # untimed steps before the warm-up: one GPU captures the step's CUDA graphs (3 rotation phases) during them
    e = case.make_engine(steps=PRIME + 2 * W + 2 * K, x0=x0, nxl=nxl, dtype=dtype, arith=arith, device=local, kernel=args.kernel)
    halo = "none"
    if world > 1:
        halo = e.connect(rank, world, allgather, broadcast)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- pass 1: `value` on the production launch path (no per-launch events; graph replay on one GPU) ----
    e.run(PRIME)
    e.run(W)
    e.sync()
    l0 = e.launch_count
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    ms = e.run_timed(K)
    barrier()
    clocks = sampler.finish() if sampler else None
    launches = e.launch_count - l0
    # ---- pass 2: the dominant kernel's own time (event pair around every stencil launch) ----------------
    e.profile(1)
    e.run(W)
    e.sync()
    e.profile(1)
    barrier()
    ms_prof = e.run_timed(K)
    barrier()
    kms, kn = e.profile(0)
    if dist is not None:
        import torch
        t = torch.tensor([ms, kms, ms_prof], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, kms, ms_prof = float(t[0]), float(t[1]), float(t[2])
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
        fields = tuple(args.e2e_fields.split(","))
        e2e_steps = max(K, args.e2e_steps)
        every = args.e2e_record_every or (1 if e2e_steps <= 2000 else -(-e2e_steps // 1000))     # long runs: <= 1000 frames
        e2e = run_e2e(case, n, rank, local, K, dtype, arith, dist, fields=fields, steps=e2e_steps, record_every=every)
        if n == 1 and e2e["file_dir"] != args.disk_dir and not args.no_disk:
            # the same run with the output on the temp directory's file system (the default above is tmpfs)
            try:
                d2 = run_e2e(case, n, rank, local, K, dtype, arith, dist, fields=fields, steps=e2e_steps, out_dir=args.disk_dir, record_every=every)
                e2e["disk"] = {k: d2[k] for k in ("value", "file_dir", "run_ms", "init_s", "writer_write_ms", "writer_wait_ms", "steps")}
            except Exception as exc:
                e2e["disk"] = {"error": str(exc)[:200]}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = hbm_peak()
    cells_local = nxl * NY * NZ
    # the stencil launches of the K profiled steps process cells_local * K cell updates in total (with slabs on the
    # NCCL path a step is an edge launch per neighbour + one interior launch); kms is their summed device time on the
    # slowest rank
    ach = b_alg(dtype) * cells_local * K / (kms * 1e-3) / 1e9 if kms > 0 else None     # GB/s of one GPU's kernel
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic = tj.get("%s_512" % dtype, {}).get("dram_bytes_per_launch") if nxl == 512 else None
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": K, "warmup": W, "ms_per_step": ms / K,
        "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": {"workload": workload_name(n, args.scaling),
                   "arith": arith, "material": "indexed (1-byte stencil class)", "kernel": info["kernel"],
                   "l2": "inputs (%.1f GB/GPU) far exceed the 126 MB L2; no explicit flush" % (info["device_bytes"] / 1e9),
                   "launch_path": "CUDA-graph replay of the step" if (n == 1 or halo in ("p2p", "fused")) else "stream launches",
                   "halo": {"none": "none", "nccl": "NCCL send/recv of 3 planes per direction per step, overlapped with the interior update",
                            "p2p": "peer memory: after the faces of a step one kernel stores the two finished edge planes into the neighbours' ghost "
                                   "planes over NVLink (CUDA IPC) and publishes the step number; a 1-thread kernel waits for the neighbours'; no collective call; "
                                   "the whole slab step is replayed as a CUDA graph",
                            "fused": "the stencil kernel stores its edge planes into the neighbours' ghost planes over NVLink (CUDA IPC) while it computes them, "
                                     "flag write/wait kernels, no collective call"}[halo]},
        "clocks": clocks, "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": (ach / peak) if ach else None,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "k_step_march", "kernel_ms_per_step": kms / K if K else None, "kernel_launches_per_step": kn / K if K else None,
                     "algorithmic_bytes_per_cell": b_alg(dtype), "kernel_share_of_step": kms / ms_prof if ms_prof else None,
                     "how": "second pass of K steps with a CUDA event pair around every stencil launch (ms_per_step there: %.4f); "
                            "`value` is timed without them" % (ms_prof / K),
                     "step_frac": b_alg(dtype) * cells_local * K / (ms * 1e-3) / 1e9 / peak},
    }
    if parity is not None:
        line["parity"] = parity
    if e2e is not None:
        line["e2e"] = e2e
    if n == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline_leg()
    if args.extra:
        line["extra"] = extra_runs(case, local)
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def cpu_baseline_leg():
    """cpu_baseline of the b200 arm (rank 0, N = 1): the unmodified reference on a bounded sample (~15 s with the
    mesh / material build), else the NumPy port."""
    try:
        ref = reference_sample(5)
    except Exception:
        ref = None
    if ref is not None:
        runs, host, _root = ref
        r, d = runs["solver_threading"], runs["solver_default"]
        return {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "reference",
                "sample": "128^3 block of the same crystal, 5 steps, UNMODIFIED reference solver_threading (6 worker threads) via "
                          "Solver.init/run; solver_default (1 thread) on the same block: %.5f Gcell/s" % d["value"],
                "host": host}
    v, sample = cpu_sample(1, 5, 1)
    return {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample + " (oracle/_ref absent)"}


def run_e2e(case, n, rank, local, K, dtype, arith, dist, fields=("ux", "uy", "uz"), steps=200, out_dir=None, record_every=1):
    """Solver.init + Solver.run through the plugin with host buffers: mesh lines, inclusion list
    and tables go host->device in init(); inside the timed run() every chunk's source samples go
    host->device and every step's surface planes come back through the pinned ring into the HDF5
    file (native writer threads).  Timed region = run() (the reference times its loop the same way,
    base_solver.py:239-279); init() is reported separately (`init_s`)."""
    import tempfile
    from phonomena_b200.solver_b200 import Solver
    g, m = case.as_grid_material()
    s = Solver()
    s.cfg.update({"precision": {"f64": "fp64", "f32": "fp32"}[dtype], "arith": arith, "device": local, "wave": "sin",
                  "wave_args": {"f": 100}, "write_mode": "thread", "record": "surface", "record_every": int(record_every),
                  "slabs_from_env": n > 1, "chunk_steps": 25,
                  "record_fields": list(fields),
                  "merge_slabs": False})      # N > 1: one slab file per rank (concatenating them is post-processing)
    # output file: tmpfs when the box has one with room (the run measures the solver and its copies, not the
    # scratch disk of the box), else the temp directory; reported in e2e["file_dir"]; N = 1 repeats the run on
    # the temp directory (e2e["disk"])
    nx, ny, nz = case.shape
    frame_bytes = 8 * sum(v for k, v in (("ux", (nx - 1) * ny), ("uy", nx * (ny - 1)), ("uz", nx * ny)) if k in fields) // n
    need = frame_bytes * (steps // record_every) + 9 * nx * ny * nz // n + (1 << 28)
    if out_dir is None:
        out_dir = tempfile.gettempdir()
        try:
            st = os.statvfs("/dev/shm")
            if st.f_bavail * st.f_frsize > 2 * need * max(1, n):
                out_dir = "/dev/shm"
        except OSError:
            pass
    s.file = os.path.join(out_dir, "phb_bench_rank%d.h5" % rank)
    t_i = time.perf_counter()
    s.init(g, m, steps)
    init_s = time.perf_counter() - t_i
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
    path = s.file if n == 1 else "%s.rank%d" % (s.file, rank)
    out = {"value": nx * ny * nz * steps / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": 8,
           "d2h_bytes_per_step": 8 * sum(v for k, v in (("ux", (nx - 1) * ny), ("uy", nx * (ny - 1)), ("uz", nx * ny)) if k in fields),
           "steps": steps, "record_every": int(record_every),
           "what": "Solver.run(): source sample H2D + surface %s planes D2H (pinned ring) -> native writer threads -> HDF5 every %s" % (
                   ",".join(fields), "step" if record_every == 1 else "%d steps (long run: the byte counts are per recorded step)" % record_every)
                   + ("; one slab file per rank, max over ranks" if n > 1 else ""),
           "file_bytes": os.path.getsize(path), "file_dir": out_dir, "init_s": init_s,
           "frames_written": int(s.writer.written) if s.writer is not None else 0,
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 512 planes per GPU (config #3 / #5); strong: the fixed 1024x512x512 grid of config #4 split over the GPUs")
    ap.add_argument("--e2e-steps", type=int, default=200, help="the e2e run() is timed over max(K, this) steps")
    ap.add_argument("--e2e-record-every", type=int, default=0, help="0 = every step up to 2000 e2e steps, else so that <= 1000 frames are written")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the slab-vs-single-GPU check before the timed region")
    ap.add_argument("--no-disk", action="store_true", help="N = 1: skip the second e2e run with the output on --disk-dir")
    ap.add_argument("--disk-dir", default=__import__("tempfile").gettempdir())
    ap.add_argument("--quick-reference", action="store_true", help="--impl reference: skip the 256^3 runs")
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
