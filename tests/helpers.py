"""Shared helpers for the parity tests (CPU side)."""
import glob
import json
import os

import numpy as np

from oracle import fdtd_numpy as onp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """Solver fixtures (the spectrum_* fixtures have their own layout, tests/test_spectrum_cpu.py)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith(("spectrum_", "periodic_y_"))]


def periodic_names():
    """Fixtures made with the reference's archived periodic-y stubs switched on (oracle/gen_golden.periodic_y_solver)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n.startswith("periodic_y_")]


def load_golden(name):
    d = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    d["wave"] = str(d["wave"])
    d["wave_args"] = json.loads(str(d["wave_args"]))
    d["steps"] = int(d["steps"])
    d["dt"] = float(d["dt"])
    d["courant"] = float(d["courant"])
    d["prim_p"] = float(d["prim_p"])
    d["sec_p"] = float(d["sec_p"])
    return d


def targets_of(d):
    return onp.make_targets(d["targets"].tolist()) if d["targets"].size else onp.make_targets([])


def oracle_from_golden(d, threads=1, **kw):
    """Oracle solver built ONLY from the fixture's inputs (mesh lines, inclusions, tables)."""
    t = targets_of(d)
    C, P = onp.set_constants(d["x"], d["y"], d["z"], t, d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    fdx, fdy, fdz, *_ = onp.spacings(d["x"], d["y"], d["z"])
    dt = onp.cfl_time_step(fdx, fdy, fdz, d["courant"], d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    return onp.OracleSolver(d["x"], d["y"], d["z"], C, P, dt, wave=d["wave"], wave_args=d["wave_args"], threads=threads, **kw)


def rel_l2(a, b):
    """Combined relative L2 over a list of arrays (SURVEY 8d parity metric)."""
    num = sum(float(np.sum((np.asarray(x, np.float64) - np.asarray(y, np.float64)) ** 2)) for x, y in zip(a, b))
    den = sum(float(np.sum(np.asarray(y, np.float64) ** 2)) for y in b)
    return (num / den) ** 0.5 if den > 0 else num ** 0.5


# ------------------------------------------------------------------------------------------
# GPU side: build a device engine from a fixture's INPUTS only (never from its outputs)
# ------------------------------------------------------------------------------------------
def engine_from_case(x, y, z, targets, prim_c, prim_p, sec_c, sec_p, courant, wave, wave_args, steps,
                     dtype="f64", arith="exact", kernel="auto", ids=None, **kw):
    from phonomena_b200 import _lib, hostmath as hm
    fdx, fdy, fdz, sdx, sdy, sdz = hm.spacings(x, y, z)
    prim, sec = {"c": prim_c, "p": prim_p}, {"c": sec_c, "p": sec_p}
    dt = hm.cfl_dt(fdx, fdy, fdz, courant, prim, sec)
    e = _lib.Engine(len(x), len(y), len(z), dt, dtype=dtype, arith=arith, kernel=kernel, **kw)
    e.set_spacing(fdx, fdy, fdz, sdx, sdy, sdz)
    e.set_material_table([prim_c, sec_c], [prim_p, sec_p])
    if ids is None:
        e.gen_material_ids(np.asarray(targets, np.float32).reshape(-1, 4), x, y, z)
        corner = int(e.get_material_ids()[0, 0, 0]) if e.x0 == 0 else 0
    else:
        e.set_material_ids(ids[e.x0:e.x0 + e.id_planes()])
        corner = int(ids[0, 0, 0])
    cc, cp = (sec_c, sec_p) if corner else (prim_c, prim_p)
    e.set_abc(hm.abc_coefficients(cc, cp, dt, fdx, fdy, fdz, sdx, sdy, sdz))
    if wave is not None:
        e.set_source_table(hm.source_table(wave, steps, dt, wave_args))
    e.dt = dt
    return e


def engine_from_golden(d, **kw):
    return engine_from_case(d["x"], d["y"], d["z"], d["targets"], d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"],
                            d["courant"], d["wave"], d["wave_args"], d["steps"], **kw)
