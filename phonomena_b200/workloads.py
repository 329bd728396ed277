"""Synthetic workloads of BASELINE.json (SURVEY 8d "Synthetic inputs"): homogeneous GaAs
blocks and phononic crystals -- a square lattice of full-depth Au cylinders (pitch 32, r = 8,
centres at 16 + 32 m) in GaAs on the uniform integer mesh the reference's buildMesh yields
for this lattice.  Material tables are those of the reference's data/default.json."""
from __future__ import annotations

import numpy as np

from . import _lib, hostmath as hm

# data/default.json "material.properties" (unscaled; Material.init multiplies by 1e10)
PROPS = {
    "GaAs": {"name": "Gallium Arsenide", "p": 5307,
             "c": [[11.88, 5.87, 5.38, 0, 0, 0], [5.87, 11.88, 5.38, 0, 0, 0], [5.87, 5.38, 11.88, 0, 0, 0],
                   [0, 0, 0, 5.94, 0, 0], [0, 0, 0, 0, 5.94, 0], [0, 0, 0, 0, 0, 5.94]]},
    "Au": {"name": "Gold", "p": 19300,
           "c": [[19.25, 16.3, 16.3, 0, 0, 0], [16.3, 19.25, 16.3, 0, 0, 0], [16.3, 16.3, 19.25, 0, 0, 0],
                 [0, 0, 0, 4.24, 0, 0], [0, 0, 0, 0, 4.24, 0], [0, 0, 0, 0, 0, 4.24]]},
}


def scaled(name):
    return np.array(PROPS[name]["c"], np.float64) * 1e10, float(PROPS[name]["p"])


class Case:
    """Inputs of one run: mesh lines, inclusion list, two materials, courant, source."""

    def __init__(self, x, y, z, targets, primary="GaAs", secondary="Au", courant=0.1, wave="sin", wave_args=None):
        self.x, self.y, self.z = (np.asarray(a, np.float64) for a in (x, y, z))
        self.targets = np.asarray(targets, np.float32).reshape(-1, 4)
        self.prim_c, self.prim_p = scaled(primary)
        self.sec_c, self.sec_p = scaled(secondary)
        self.names = (PROPS[primary]["name"], PROPS[secondary]["name"])
        self.courant, self.wave, self.wave_args = courant, wave, dict(wave_args or {"f": 100})
        self.sp = hm.spacings(self.x, self.y, self.z)
        self.dt = hm.cfl_dt(self.sp[0], self.sp[1], self.sp[2], courant, {"c": self.prim_c, "p": self.prim_p},
                            {"c": self.sec_c, "p": self.sec_p})

    @property
    def shape(self):
        return (self.x.size, self.y.size, self.z.size)

    def as_grid_material(self):
        """Duck-typed stand-ins for the reference's Grid / Material objects carrying exactly the
        attributes Solver.init reads (mesh lines, float32 inclusion records, SI_conversion;
        primary / secondary with the already-scaled 6x6 table, density and name; c_max; grid)."""
        from types import SimpleNamespace
        tdt = np.dtype([("x", "f"), ("y", "f"), ("z", "f"), ("r", "f")])      # grid.py:39
        t = np.array([tuple(r) for r in self.targets], dtype=tdt).reshape(-1)
        g = SimpleNamespace(x=self.x.copy(), y=self.y.copy(), z=self.z.copy(), targets=t, SI_conversion=1)
        m = SimpleNamespace(primary={"c": self.prim_c.copy(), "p": self.prim_p, "name": self.names[0]},
                            secondary={"c": self.sec_c.copy(), "p": self.sec_p, "name": self.names[1]},
                            c_max=self.courant, grid=g)
        return g, m

    def make_engine(self, steps, x0=0, nxl=None, source_start=0, **kw):
        nx, ny, nz = self.shape
        e = _lib.Engine(nx, ny, nz, self.dt, x0=x0, nxl=nxl, **kw)
        e.set_spacing(*self.sp)
        e.set_material_table([self.prim_c, self.sec_c], [self.prim_p, self.sec_p])
        e.gen_material_ids(self.targets, self.x, self.y, self.z)
        # corner cell (0,0,0) of the lattice cases is always primary (centres >= 16, r = 8)
        e.set_abc(hm.abc_coefficients(self.prim_c, self.prim_p, self.dt, *self.sp))
        if self.wave is not None:
            e.set_source_table(hm.source_table(self.wave, steps, self.dt, self.wave_args, start=source_start))
        return e


def lattice_targets(nx, ny, nz, pitch=32.0, r=8.0):
    """Centres 16 + 32 m with x +- r strictly inside the domain (grid.py:151-152)."""
    sx, sy, sz = nx - 1, ny - 1, nz - 1
    out = []
    cx = pitch / 2
    while cx + r < sx:
        cy = pitch / 2
        while cy + r < sy:
            out.append((cx, cy, float(sz), r))
            cy += pitch
        cx += pitch
    return np.array(out, np.float32).reshape(-1, 4)


def crystal_case(nx, ny, nz, homogeneous=False, **kw):
    x, y, z = np.arange(nx, dtype=np.float64), np.arange(ny, dtype=np.float64), np.arange(nz, dtype=np.float64)
    t = np.zeros((0, 4), np.float32) if homogeneous else lattice_targets(nx, ny, nz)
    return Case(x, y, z, t, secondary="GaAs" if homogeneous else "Au", **kw)
