"""Host-side scalar/1-D preparation for the device engine.

Everything here is O(Nx+Ny+Nz) or O(steps) work the reference also does on the host in
NumPy; it is evaluated with the same expressions so the inputs handed to the device are
bit-identical to the reference's (the per-cell work all happens on the GPU).
"""
from __future__ import annotations

import numpy as np

F64 = np.float64


def spacings(x, y, z, si_conversion=1):
    """fd* = diff of mesh lines, sd* = mean of adjacent fd  (grid.py:112-127).
    Returned flat (1-D)."""
    xs, ys, zs = (np.asarray(a, F64) * si_conversion for a in (x, y, z))
    fdx, fdy, fdz = xs[1:] - xs[:-1], ys[1:] - ys[:-1], zs[1:] - zs[:-1]
    sd = lambda f: np.mean([f[1:], f[:-1]], axis=0)
    return fdx, fdy, fdz, sd(fdx), sd(fdy), sd(fdz)


def cfl_dt(fdx, fdy, fdz, courant, prim, sec, si_conversion=1):
    """material.py:80-93.  prim/sec: dicts with scaled 'c' (6x6) and 'p'."""
    def one(c, p):
        vl = np.sqrt(c[0][0] / p)
        vt = np.sqrt(c[3][3] / p)
        vmax = max((vl, vt))
        dxmin = min((np.amin(fdx), np.amin(fdy), np.amin(fdz))) * si_conversion
        return courant * dxmin / vmax
    return min((one(prim["c"], prim["p"]), one(sec["c"], sec["p"])))


def abc_coefficients(c_corner, p_corner, dt, fdx, fdy, fdz, sdx, sdy, sdz):
    """Mur coefficients of apply_u_abc (base_solver.py:525-537) for the corner cell's
    6x6 stiffness and density.  Returns the eight scalars in the engine's order."""
    c11, c44 = c_corner[0][0], c_corner[3][3]
    vl = np.sqrt(c11 / p_corner)
    vt = np.sqrt(c44 / p_corner)
    k = lambda v, d: (v * dt - d) / (v * dt + d)
    return {
        "clx": float(k(vl, np.asarray(sdx, F64))[-1]), "ctx": float(k(vt, np.asarray(fdx, F64))[-1]),
        "cly0": float(k(vl, np.asarray(sdy, F64))[0]), "cty0": float(k(vt, np.asarray(fdy, F64))[0]),
        "cly1": float(k(vl, np.asarray(sdy, F64))[-1]), "cty1": float(k(vt, np.asarray(fdy, F64))[-1]),
        "clz": float(k(vl, np.asarray(sdz, F64))[-1]), "ctz": float(k(vt, np.asarray(fdz, F64))[-1]),
    }


def wave_sin(tt, dt, f, **_):
    """base_solver.py:294-299"""
    return np.sin(2 * np.pi * f * tt * dt)


def wave_ricker(tt, dt, f, source_delay=0, **_):
    """base_solver.py:301-312"""
    arg = (np.pi * f * (dt * tt - source_delay)) ** 2
    return (1 - 2 * arg) * np.exp(-arg)


WAVES = {"sin": wave_sin, "ricker": wave_ricker}


def source_table(kind, steps, dt, wave_args, start=0):
    """w(tt) for tt = start .. start+steps-1, one scalar evaluation per step as in the
    reference loop (base_solver.py:251)."""
    fn = WAVES[kind]
    return np.array([fn(tt=tt, dt=dt, **wave_args) for tt in range(start, start + steps)], F64)


def nonlinspace(spacing):
    """simulation/analysis.py:9-18: positions from a spacing array (the first spacing is skipped)."""
    spacing = np.asarray(spacing, F64).reshape(-1)
    X = np.zeros(spacing.size)
    for i in range(1, spacing.size):
        X[i] = X[i - 1] + spacing[i]
    return X


def split_slabs(nx, nparts, min_planes=4):
    """Contiguous x-slabs [x0, x0+n) as evenly as possible."""
    if nparts < 1 or nx < nparts * min_planes:
        raise ValueError("cannot split %d planes into %d slabs of >= %d planes" % (nx, nparts, min_planes))
    base, rem = divmod(nx, nparts)
    out, x0 = [], 0
    for r in range(nparts):
        n = base + (1 if r < rem else 0)
        out.append((x0, n))
        x0 += n
    return out


def halo_plan(rank, nranks, nxl):
    """Halo exchange of one x-slab with its neighbours, as (peer, local plane sent, local plane
    received into); local plane l = i - x0 + 1, so 1 and nxl are the first / last owned planes and
    0 / nxl + 1 the ghosts (phb200.cu exchange(): same pairs, three components each)."""
    plan = []
    if rank > 0:
        plan.append((rank - 1, 1, 0))
    if rank < nranks - 1:
        plan.append((rank + 1, nxl, nxl + 1))
    return plan
