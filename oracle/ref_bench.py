"""Time the UNMODIFIED reference solvers on the host CPU -- TEST / MEASUREMENT INFRASTRUCTURE ONLY.

Used by ``bench.py --impl reference`` and by the ``cpu_baseline`` leg of the b200 arm (the two places
the bench may execute ``oracle/``).  The reference is imported through ``oracle/refshim.py`` from
``/root/reference`` (build container) or from the copy ``make -C oracle ref`` staged in ``oracle/_ref/``
(GPU box); nothing of it is edited.  The run goes through the reference's own public path
(SURVEY 8d "CPU baseline timing"):

    common.importSolver("solver_default" | "solver_threading")      common.py:88-91
    solver.cfg["write_mode"] = "off"                                 tests/test_speed.py:47-56
    solver.init(grid, material, steps)                               base_solver.py:194-222
    solver.run()                                                     base_solver.py:224-280

on a grid / material built by the reference's own ``Grid`` / ``Material`` classes: an n^3 block of the
bench's phononic crystal (Au cylinders pitch 32, r 8, full depth, in GaAs; uniform integer mesh).
The rate is cells * steps / (wall time of run()), the reference's own clock (base_solver.py:239,278).
"""
import json
import os
import platform
import time

import numpy as np

from oracle import refshim


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return platform.processor() or "unknown"


def host_info():
    return {"cpu_model": cpu_model(), "cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
            "numpy": np.__version__, "python": platform.python_version()}


def build_case(n, homogeneous=False, pitch=32.0, r=8.0):
    """Grid + Material of the reference for an n^3-point block of the bench crystal (size n-1, unit spacing)."""
    refshim.install()
    from simulation import grid as rgrid, material as rmat
    props = json.load(open(os.path.join(refshim.REF_ROOT, "data", "default.json")))["material"]["properties"]
    g = rgrid.Grid()
    g.init(size_x=n - 1, size_y=n - 1, size_z=n - 1)
    g.min_d = 1
    g.max_dx = g.max_dy = g.max_dz = 1
    g.slope = 1.0
    if not homogeneous:
        c = pitch / 2
        centres = []
        while c + r < n - 1:
            centres.append(c)
            c += pitch
        for cx in centres:
            for cy in centres:
                g.addInclusion(x=cx, y=cy, z=float(n - 1), r=r)
    g.buildMesh()
    g.update()
    m = rmat.Material()
    m.init(grid=g, properties=props)
    m.c_max = 0.1
    m.setPrimary("GaAs")
    m.setSecondary("GaAs" if homogeneous else "Au")
    m.update()
    return g, m


def time_solver(module, g, m, steps, warmup=0, cfg=None):
    """One reference solver through init() + run(); returns (Gcell/s, seconds of run(), final |uz|)."""
    common = refshim.install()
    s = common.importSolver(module)
    s.cfg["write_mode"] = "off"
    s.cfg.update({"wave": "sin", "wave_args": {"f": 100}})
    s.cfg.update(cfg or {})
    if warmup:
        s.init(g, m, warmup)
        s.run()
    s.init(g, m, steps)
    t0 = time.perf_counter()
    s.run()
    dt = time.perf_counter() - t0
    cells = g.x.size * g.y.size * g.z.size
    return cells * steps / dt / 1e9, dt, float(np.linalg.norm(s.g.uz))


def run(n=128, steps=10, warmup=1, solvers=("solver_threading", "solver_default"), homogeneous=False):
    g, m = build_case(n, homogeneous)
    out = {}
    for mod in solvers:
        v, dt, nrm = time_solver(mod, g, m, steps, warmup)
        out[mod] = {"value": v, "unit": "Gcell/s", "seconds": dt, "steps": steps, "grid": [int(g.x.size), int(g.y.size), int(g.z.size)],
                    "uz_norm": nrm, "threads": 6 if mod == "solver_threading" else 1}
    return out


if __name__ == "__main__":
    import sys
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    print(json.dumps({"host": host_info(), "ref_root": refshim.REF_ROOT, "runs": run(n, k)}, indent=1))
