"""Shared helpers for the parity tests (CPU side)."""
import glob
import json
import os

import numpy as np

from oracle import fdtd_numpy as onp

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


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


def oracle_from_golden(d, threads=1):
    """Oracle solver built ONLY from the fixture's inputs (mesh lines, inclusions, tables)."""
    t = targets_of(d)
    C, P = onp.set_constants(d["x"], d["y"], d["z"], t, d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    fdx, fdy, fdz, *_ = onp.spacings(d["x"], d["y"], d["z"])
    dt = onp.cfl_time_step(fdx, fdy, fdz, d["courant"], d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    return onp.OracleSolver(d["x"], d["y"], d["z"], C, P, dt, wave=d["wave"], wave_args=d["wave_args"], threads=threads)


def rel_l2(a, b):
    """Combined relative L2 over a list of arrays (SURVEY 8d parity metric)."""
    num = sum(float(np.sum((np.asarray(x, np.float64) - np.asarray(y, np.float64)) ** 2)) for x, y in zip(a, b))
    den = sum(float(np.sum(np.asarray(y, np.float64) ** 2)) for y in b)
    return (num / den) ** 0.5 if den > 0 else num ** 0.5
