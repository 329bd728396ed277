"""CPU oracle for the Phonomena FDTD time-stepping path  --  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference algorithm.  It is the *checker*
for the CUDA path; nothing under ``phonomena_b200/`` imports it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it.

Parity status: PINNED.  ``tests/test_oracle_golden.py`` checks this restatement
bit-for-bit (``np.array_equal``) against field dumps produced by the unmodified
reference solver (``solver_default``) in the build container -- see
``oracle/gen_golden.py`` (generator) and ``tests/golden/*.npz`` (fixtures) -- and
against the known answers in SURVEY.md App. D.  The reference has no golden
vectors of its own (its tests only check "no exception").

Every function names the reference lines it restates (paths relative to the
reference checkout, ``phonomena/simulation/...``).  Array conventions are the
reference's: C order, z fastest; ``ux (Nx-1,Ny,Nz)``, ``uy (Nx,Ny-1,Nz)``,
``uz (Nx,Ny,Nz-1)``; ``T1..T3 (Nx,Ny,Nz)``, ``T4 (Nx,Ny-1,Nz-1)``,
``T5 (Nx-1,Ny,Nz-1)``, ``T6 (Nx-1,Ny-1,Nz)`` (grid.py:88-110); float64 only.

All arithmetic keeps the reference's expression order (``C*diff/sd`` evaluated
left to right, sums left to right) so that results are bit-identical to the
reference's NumPy evaluation, not merely close.
"""
from __future__ import annotations

from concurrent.futures import ThreadPoolExecutor

import numpy as np

F64 = np.float64
# dtype of an inclusion record: four float32 fields (grid.py:39)
TARGET_DTYPE = np.dtype([("x", "f"), ("y", "f"), ("z", "f"), ("r", "f")])


# ----------------------------------------------------------------------------
# Mesh spacing, material indexing, time step  (input contract of the path)
# ----------------------------------------------------------------------------
def spacings(x, y, z):
    """Full (fd*) and staggered (sd*) spacings, shaped for broadcasting.

    Restates grid.py:118-127.  ``np.mean([a, b], axis=0)`` there equals
    ``(a + b) / 2`` evaluated as add-then-divide, which is what is done here.
    """
    x = np.asarray(x, F64)
    y = np.asarray(y, F64)
    z = np.asarray(z, F64)
    fdx = (x[1:] - x[:-1]).reshape(-1, 1, 1)
    fdy = (y[1:] - y[:-1]).reshape(1, -1, 1)
    fdz = (z[1:] - z[:-1]).reshape(1, 1, -1)
    sdx = (fdx[1:, :, :] + fdx[:-1, :, :]) / 2.0
    sdy = (fdy[:, 1:, :] + fdy[:, :-1, :]) / 2.0
    sdz = (fdz[:, :, 1:] + fdz[:, :, :-1]) / 2.0
    return fdx, fdy, fdz, sdx, sdy, sdz


def make_targets(rows):
    """Inclusion list as the float32 record array of grid.py:39,144-156.

    ``rows`` is an iterable of (x, y, z, r)."""
    rows = [tuple(float(v) for v in r) for r in rows]
    return np.array(rows, dtype=TARGET_DTYPE).reshape(-1)


def inclusion_indices(x, y, z, targets):
    """(y,x) index pairs and z indices per inclusion.  Restates grid.py:158-175.

    Quirk kept on purpose: ``z`` is re-bound to the *index array*
    after the first inclusion, so later inclusions compare indices, not
    coordinates, against ``t['z']``.  The cylinder test mixes the float64 mesh
    lines with float32 inclusion fields exactly as NumPy promotes them there.
    """
    out = []
    zz = z
    for t in targets:
        X, Y = np.meshgrid(x, y)
        X = X - t["x"]
        Y = Y - t["y"]
        R = np.sqrt(np.add(np.square(X), np.square(Y)))
        yx = np.array(np.where(R < t["r"])).transpose()
        zz = np.array(np.where(zz <= t["z"])).flatten()
        out.append((yx, zz))
    return out


def material_id_map(x, y, z, targets):
    """uint8 map (Nx,Ny,Nz): 0 = primary, 1 = secondary.  Same cells as
    material.py:55-63 writes with the secondary constants."""
    ids = np.zeros((len(x), len(y), len(z)), np.uint8)
    for yx, zi in inclusion_indices(x, y, z, targets):
        if yx.size == 0 or zi.size == 0:
            continue
        ids[yx[:, 1][:, None], yx[:, 0][:, None], zi[None, :]] = 1
    return ids


def scale_table(c):
    """Material.init multiplies the 6x6 table by 1e10 (material.py:42)."""
    return np.array(c, F64) * 1e10


def set_constants(x, y, z, targets, c_primary, p_primary, c_secondary, p_secondary):
    """Dense C (Nx,Ny,Nz,6,6) and P (Nx,Ny,Nz).  Restates material.py:48-63.
    ``c_*`` are the already-scaled 6x6 tables."""
    shape = (len(x), len(y), len(z))
    C = np.zeros(shape + (6, 6))
    P = np.zeros(shape)
    C[:, :, :] = np.array(c_primary, F64)
    P[:, :, :] = float(p_primary)
    ids = material_id_map(x, y, z, targets)
    sel = ids == 1
    C[sel] = np.array(c_secondary, F64)
    P[sel] = float(p_secondary)
    return C, P


def cfl_time_step(fdx, fdy, fdz, courant, c_primary, p_primary, c_secondary, p_secondary,
                  si_conversion=1):
    """dt from the CFL condition.  Restates material.py:80-93."""
    def one(c, p):
        vl = np.sqrt(c[0][0] / p)
        vt = np.sqrt(c[3][3] / p)
        vmax = max((vl, vt))
        dmin = min((np.amin(fdx), np.amin(fdy), np.amin(fdz))) * si_conversion
        return courant * dmin / vmax
    return min((one(c_primary, p_primary), one(c_secondary, p_secondary)))


# ----------------------------------------------------------------------------
# Sources  (base_solver.py:294-312)
# ----------------------------------------------------------------------------
def source_sin(tt, dt, f, **_):
    """base_solver.py:294-299"""
    return np.sin(2 * np.pi * f * tt * dt)


def source_ricker(tt, dt, f, source_delay=0, **_):
    """base_solver.py:301-312"""
    arg = (np.pi * f * (dt * tt - source_delay)) ** 2
    return (1 - 2 * arg) * np.exp(-arg)


SOURCES = {"sin": source_sin, "ricker": source_ricker}


def source_table(kind, steps, dt, wave_args):
    """w(tt) for tt = 0..steps-1, evaluated one scalar at a time like the
    reference loop does (base_solver.py:251)."""
    fn = SOURCES[kind]
    return np.array([fn(tt=tt, dt=dt, **wave_args) for tt in range(steps)], F64)


# ----------------------------------------------------------------------------
# The stepping path
# ----------------------------------------------------------------------------
class OracleSolver:
    """State + one-step update restating base_solver.py:245-260,323-571.

    Parameters are plain arrays: mesh lines, dense ``C``/``P`` in the reference
    layout, ``dt``.  ``threads`` > 1 runs the six stress updates and the three
    displacement updates as separate tasks on a thread pool -- the scheme of the
    reference's ``solver_threading.py:42-102`` (bit-identical results: the tasks
    write disjoint arrays).
    """

    def __init__(self, x, y, z, C, P, dt, wave="sin", wave_args=None, threads=1, bc_y="absorbing"):
        assert bc_y in ("absorbing", "periodic")
        self.bc_y = bc_y
        self.x = np.asarray(x, F64)
        self.y = np.asarray(y, F64)
        self.z = np.asarray(z, F64)
        nx, ny, nz = self.x.size, self.y.size, self.z.size
        self.shape = (nx, ny, nz)
        self.C = C
        self.P = P
        self.dt = dt
        self.wave = wave
        self.wave_args = dict(wave_args or {"f": 100})
        self.fdx, self.fdy, self.fdz, self.sdx, self.sdy, self.sdz = spacings(x, y, z)
        z3 = lambda *s: np.zeros(s, F64)
        # grid.py:88-110
        self.ux, self.uy, self.uz = z3(nx - 1, ny, nz), z3(nx, ny - 1, nz), z3(nx, ny, nz - 1)
        self.ux_new, self.uy_new, self.uz_new = z3(nx - 1, ny, nz), z3(nx, ny - 1, nz), z3(nx, ny, nz - 1)
        self.ux_old, self.uy_old, self.uz_old = z3(nx - 1, ny, nz), z3(nx, ny - 1, nz), z3(nx, ny, nz - 1)
        self.T1, self.T2, self.T3 = z3(nx, ny, nz), z3(nx, ny, nz), z3(nx, ny, nz)
        self.T4, self.T5, self.T6 = z3(nx, ny - 1, nz - 1), z3(nx - 1, ny, nz - 1), z3(nx - 1, ny - 1, nz)
        self.tt = 0
        self._pool = ThreadPoolExecutor(threads) if threads > 1 else None
        self.threads = threads

    # -- helpers -----------------------------------------------------------
    def _parallel(self, fns):
        if self._pool is None:
            for f in fns:
                f()
        else:
            for fut in [self._pool.submit(f) for f in fns]:
                fut.result()

    # -- A.1  update_T  (base_solver.py:323-372) ----------------------------
    def _normal(self, T, row):
        C, ux, uy, uz = self.C, self.ux, self.uy, self.uz
        I = (slice(1, -1),) * 3
        T[I] = (
            C[1:-1, 1:-1, 1:-1, row, 0] * (ux[1:, 1:-1, 1:-1] - ux[:-1, 1:-1, 1:-1]) / self.sdx
            + C[1:-1, 1:-1, 1:-1, row, 1] * (uy[1:-1, 1:, 1:-1] - uy[1:-1, :-1, 1:-1]) / self.sdy
            + C[1:-1, 1:-1, 1:-1, row, 2] * (uz[1:-1, 1:-1, 1:] - uz[1:-1, 1:-1, :-1]) / self.sdz
        )

    def _T4(self):
        C, uy, uz = self.C, self.uy, self.uz
        self.T4[1:-1, :, :] = C[1:-1, 1:, 1:, 3, 3] * (
            (uy[1:-1, :, 1:] - uy[1:-1, :, :-1]) / self.fdz
            + (uz[1:-1, 1:, :] - uz[1:-1, :-1, :]) / self.fdy
        )

    def _T5(self):
        C, ux, uz = self.C, self.ux, self.uz
        self.T5[:, 1:-1, :] = C[1:, 1:-1, 1:, 4, 4] * (
            (ux[:, 1:-1, 1:] - ux[:, 1:-1, :-1]) / self.fdz
            + (uz[1:, 1:-1, :] - uz[:-1, 1:-1, :]) / self.fdx
        )

    def _T6(self):
        C, ux, uy = self.C, self.ux, self.uy
        self.T6[:, :, 1:-1] = C[1:, 1:, 1:-1, 5, 5] * (
            (ux[:, 1:, 1:-1] - ux[:, :-1, 1:-1]) / self.fdy
            + (uy[1:, :, 1:-1] - uy[:-1, :, 1:-1]) / self.fdx
        )

    def update_T(self):
        self._parallel([
            lambda: self._normal(self.T1, 0),
            lambda: self._normal(self.T2, 1),
            lambda: self._normal(self.T3, 2),
            self._T4, self._T5, self._T6,
        ])

    # -- A.2  apply_T_tfbc  (base_solver.py:402-433) -------------------------
    def apply_T_tfbc(self):
        C, ux, uy, uz = self.C, self.ux, self.uy, self.uz
        # first-element spacings only (1x1 arrays in the reference)
        sdx0, sdy0, sdz0 = self.sdx[0, :, :], self.sdy[:, 0, :], self.sdz[:, :, 0]
        fdx0, fdy0, fdz0 = self.fdx[0, :, :], self.fdy[:, 0, :], self.fdz[:, :, 0]
        for T, row in ((self.T1, 0), (self.T2, 1)):
            T[1:-1, 1:-1, 0] = (
                C[1:-1, 1:-1, 0, row, 0] * (ux[1:, 1:-1, 0] - ux[:-1, 1:-1, 0]) / sdx0
                + C[1:-1, 1:-1, 0, row, 1] * (uy[1:-1, 1:, 0] - uy[1:-1, :-1, 0]) / sdy0
                + C[1:-1, 1:-1, 0, row, 2] * (uz[1:-1, 1:-1, 0] - 0) / sdz0
            )
        self.T3[1:-1, 1:-1, 0] = 0
        # note the "wrong-axis" spacings -- kept as in the reference
        self.T4[1:-1, :, 0] = C[1:-1, 1:, 0, 3, 3] * (
            (uy[1:-1, :, 1] - uy[1:-1, :, 0]) / fdy0 + (uz[1:-1, 1:, 0] - uz[1:-1, :-1, 0]) / fdz0
        )
        self.T5[:, 1:-1, 0] = C[1:, 1:-1, 0, 4, 4] * (
            (ux[:, 1:-1, 1] - ux[:, 1:-1, 0]) / fdx0 + (uz[1:, 1:-1, 0] - uz[:-1, 1:-1, 0]) / fdz0
        )
        self.T6[:, :, 0] = C[1:, 1:, 0, 5, 5] * (
            (ux[:, 1:, 0] - ux[:, :-1, 0]) / fdx0 + (uy[1:, :, 0] - uy[:-1, :, 0]) / fdz0
        )

    # -- A.3  update_u  (base_solver.py:435-463) -----------------------------
    def _ux(self):
        d2 = self.dt ** 2
        T1, T5, T6 = self.T1, self.T5, self.T6
        self.ux_new[:, 1:-1, 1:-1] = (
            2 * self.ux[:, 1:-1, 1:-1] - self.ux_old[:, 1:-1, 1:-1]
            + (d2 / self.P[1:, 1:-1, 1:-1]) * (
                (T1[1:, 1:-1, 1:-1] - T1[:-1, 1:-1, 1:-1]) / self.fdx
                + (T6[:, 1:, 1:-1] - T6[:, :-1, 1:-1]) / self.sdy
                + (T5[:, 1:-1, 1:] - T5[:, 1:-1, :-1]) / self.sdz
            )
        )

    def _uy(self):
        d2 = self.dt ** 2
        T2, T4, T6 = self.T2, self.T4, self.T6
        self.uy_new[1:-1, :, 1:-1] = (
            2 * self.uy[1:-1, :, 1:-1] - self.uy_old[1:-1, :, 1:-1]
            + (d2 / self.P[1:-1, 1:, 1:-1]) * (
                (T6[1:, :, 1:-1] - T6[:-1, :, 1:-1]) / self.sdx
                + (T2[1:-1, 1:, 1:-1] - T2[1:-1, :-1, 1:-1]) / self.fdy
                + (T4[1:-1, :, 1:] - T4[1:-1, :, :-1]) / self.sdz
            )
        )

    def _uz(self):
        d2 = self.dt ** 2
        T3, T4, T5 = self.T3, self.T4, self.T5
        self.uz_new[1:-1, 1:-1, :] = (
            2 * self.uz[1:-1, 1:-1, :] - self.uz_old[1:-1, 1:-1, :]
            + (d2 / self.P[1:-1, 1:-1, 1:]) * (
                (T5[1:, 1:-1, :] - T5[:-1, 1:-1, :]) / self.sdx
                + (T4[1:-1, 1:, :] - T4[1:-1, :-1, :]) / self.sdy
                + (T3[1:-1, 1:-1, 1:] - T3[1:-1, 1:-1, :-1]) / self.fdz
            )
        )

    def update_u(self):
        self._parallel([self._ux, self._uy, self._uz])

    # -- A.4  apply_u_tfbc  (base_solver.py:488-517) -------------------------
    def apply_u_tfbc(self):
        d2 = self.dt ** 2
        P = self.P
        T1, T2, T3, T4, T5, T6 = self.T1, self.T2, self.T3, self.T4, self.T5, self.T6
        sdx0, sdy0, sdz0 = self.sdx[0, :, :], self.sdy[:, 0, :], self.sdz[:, :, 0]
        fdx0, fdy0, fdz0 = self.fdx[0, :, :], self.fdy[:, 0, :], self.fdz[:, :, 0]
        self.ux_new[:, 1:-1, 0] = (
            2 * self.ux[:, 1:-1, 0] - self.ux_old[:, 1:-1, 0]
            + (d2 / P[1:, 1:-1, 0]) * (
                (T1[1:, 1:-1, 0] - T1[:-1, 1:-1, 0]) / fdx0
                + (T6[:, 1:, 0] - T6[:, :-1, 0]) / sdy0
                + (T5[:, 1:-1, 0] - 0) / sdz0
            )
        )
        self.uy_new[1:-1, :, 0] = (
            2 * self.uy[1:-1, :, 0] - self.uy_old[1:-1, :, 0]
            + (d2 / P[1:-1, 1:, 0]) * (
                (T6[1:, :, 0] - T6[:-1, :, 0]) / sdx0
                + (T2[1:-1, 1:, 0] - T2[1:-1, :-1, 0]) / fdy0
                + (T4[1:-1, :, 0] - 0) / sdz0
            )
        )
        # precedence slip kept: T3[..,1] is not divided; rho at k=0
        self.uz_new[1:-1, 1:-1, 0] = (
            2 * self.uz[1:-1, 1:-1, 0] - self.uz_old[1:-1, 1:-1, 0]
            + (d2 / P[1:-1, 1:-1, 0]) * (
                (T5[1:, 1:-1, 0] - T5[:-1, 1:-1, 0]) / sdx0
                + (T4[1:-1, 1:, 0] - T4[1:-1, :-1, 0]) / sdy0
                + T3[1:-1, 1:-1, 1] - T3[1:-1, 1:-1, 0] / fdz0
            )
        )

    # -- A.5  apply_u_abc  (base_solver.py:519-554) --------------------------
    def abc_coefficients(self):
        """Mur coefficients from the corner cell (base_solver.py:525-537).
        Returns dict with the eight scalars actually used."""
        c11 = self.C[0, 0, 0, 0, 0]
        c44 = self.C[0, 0, 0, 3, 3]
        vl = np.sqrt(c11 / self.P[0, 0, 0])
        vt = np.sqrt(c44 / self.P[0, 0, 0])
        dt = self.dt
        k = lambda v, d: (v * dt - d) / (v * dt + d)
        return {
            "ctx": k(vt, self.fdx)[:, 0, 0][-1], "clx": k(vl, self.sdx)[:, 0, 0][-1],
            "cty0": k(vt, self.fdy)[0, :, 0][0], "cly0": k(vl, self.sdy)[0, :, 0][0],
            "cty1": k(vt, self.fdy)[0, :, 0][-1], "cly1": k(vl, self.sdy)[0, :, 0][-1],
            "ctz": k(vt, self.fdz)[0, 0, :][-1], "clz": k(vl, self.sdz)[0, 0, :][-1],
        }

    def apply_u_abc(self):
        c = self.abc_coefficients()
        ux, uy, uz = self.ux, self.uy, self.uz
        nx_, ny_, nz_ = self.ux_new, self.uy_new, self.uz_new
        # x = -1 face
        nx_[-1, :, :] = ux[-2, :, :] + c["clx"] * (nx_[-2, :, :] - ux[-1, :, :])
        ny_[-1, :, :] = uy[-2, :, :] + c["ctx"] * (ny_[-2, :, :] - uy[-1, :, :])
        nz_[-1, :, :] = uz[-2, :, :] + c["ctx"] * (nz_[-2, :, :] - uz[-1, :, :])
        # y = 0 face
        nx_[:, 0, :] = ux[:, 1, :] + c["cty0"] * (nx_[:, 1, :] - ux[:, 0, :])
        ny_[:, 0, :] = uy[:, 1, :] + c["cly0"] * (ny_[:, 1, :] - uy[:, 0, :])
        nz_[:, 0, :] = uz[:, 1, :] + c["cty0"] * (nz_[:, 1, :] - uz[:, 0, :])
        # y = -1 face
        nx_[:, -1, :] = ux[:, -2, :] + c["cty1"] * (nx_[:, -2, :] - ux[:, -1, :])
        ny_[:, -1, :] = uy[:, -2, :] + c["cly1"] * (ny_[:, -2, :] - uy[:, -1, :])
        nz_[:, -1, :] = uz[:, -2, :] + c["cty1"] * (nz_[:, -2, :] - uz[:, -1, :])
        # z = -1 face
        nx_[:, :, -1] = ux[:, :, -2] + c["ctz"] * (nx_[:, :, -2] - ux[:, :, -1])
        ny_[:, :, -1] = uy[:, :, -2] + c["ctz"] * (ny_[:, :, -2] - uy[:, :, -1])
        nz_[:, :, -1] = uz[:, :, -2] + c["clz"] * (nz_[:, :, -2] - uz[:, :, -1])

    # -- periodic y boundaries: the reference's ARCHIVED stubs (base_solver.py:383-400, 475-486), zero Bloch phase --
    def apply_T_pbc(self):
        """base_solver.py:388-400, the uncommented lines."""
        self.T1[:, 0, :] = self.T1[:, -2, :]
        self.T2[:, 0, :] = self.T2[:, -2, :]
        self.T3[:, 0, :] = self.T3[:, -2, :]
        self.T5[:, 0, :] = self.T5[:, -2, :]
        self.T4[:, -1, :] = self.T4[:, 1, :]
        self.T6[:, -1, :] = self.T6[:, 1, :]

    def apply_u_pbc(self):
        """base_solver.py:480-486, the uncommented lines."""
        self.ux_new[:, 0, :] = self.ux_new[:, -2, :]
        self.uz_new[:, 0, :] = self.uz_new[:, -2, :]
        self.uy_new[:, -1, :] = self.uy_new[:, 1, :]

    def apply_u_abc_xz(self):
        """The x = -1 and z = -1 faces of apply_u_abc (base_solver.py:539-542, 551-554); the y faces are what the
        periodic copies replace."""
        c = self.abc_coefficients()
        ux, uy, uz = self.ux, self.uy, self.uz
        nx_, ny_, nz_ = self.ux_new, self.uy_new, self.uz_new
        nx_[-1, :, :] = ux[-2, :, :] + c["clx"] * (nx_[-2, :, :] - ux[-1, :, :])
        ny_[-1, :, :] = uy[-2, :, :] + c["ctx"] * (ny_[-2, :, :] - uy[-1, :, :])
        nz_[-1, :, :] = uz[-2, :, :] + c["ctx"] * (nz_[-2, :, :] - uz[-1, :, :])
        nx_[:, :, -1] = ux[:, :, -2] + c["ctz"] * (nx_[:, :, -2] - ux[:, :, -1])
        ny_[:, :, -1] = uy[:, :, -2] + c["ctz"] * (ny_[:, :, -2] - uy[:, :, -1])
        nz_[:, :, -1] = uz[:, :, -2] + c["clz"] * (nz_[:, :, -2] - uz[:, :, -1])

    # -- A.6  time_step  (base_solver.py:556-571) -----------------------------
    def time_step(self):
        """Copy-based shift exactly as the reference: afterwards ``u_new == u``
        (App. B #9 depends on that invariant)."""
        for a in ("ux", "uy", "uz"):
            cur, new, old = getattr(self, a), getattr(self, a + "_new"), getattr(self, a + "_old")
            old[...] = cur
            cur[...] = new

    # -- one step / run  (base_solver.py:245-260) -----------------------------
    def step(self, w=None):
        if w is None:
            w = SOURCES[self.wave](tt=self.tt, dt=self.dt, **self.wave_args)
        self.uz[0, :, 0] = w
        self.update_T()
        self.apply_T_tfbc()
        if self.bc_y == "periodic":          # each stub right after the corresponding traction-free update
            self.apply_T_pbc()
        self.update_u()
        self.apply_u_tfbc()
        if self.bc_y == "periodic":
            self.apply_u_pbc()
            self.apply_u_abc_xz()
        else:
            self.apply_u_abc()
        self.time_step()
        self.tt += 1

    def run(self, steps, on_step=None):
        for _ in range(steps):
            self.step()
            if on_step is not None:
                on_step(self)
        return self

    def close(self):
        if self._pool is not None:
            self._pool.shutdown()
            self._pool = None


def build_case(x, y, z, targets, props, primary, secondary, courant, wave="sin",
               wave_args=None, threads=1):
    """Convenience: mesh lines + inclusion list + property table -> OracleSolver,
    following loadSettings/Material.update (common.py:112-156, material.py:36-53).
    ``props[name] = {'c': 6x6 (unscaled), 'p': rho}``."""
    cp, cs = scale_table(props[primary]["c"]), scale_table(props[secondary]["c"])
    pp, ps = props[primary]["p"], props[secondary]["p"]
    fdx, fdy, fdz, *_ = spacings(x, y, z)
    dt = cfl_time_step(fdx, fdy, fdz, courant, cp, pp, cs, ps)
    C, P = set_constants(x, y, z, targets, cp, pp, cs, ps)
    return OracleSolver(x, y, z, C, P, dt, wave=wave, wave_args=wave_args, threads=threads)


class BlochOracle:
    """Bloch-periodic y boundaries with a phase, u(y + L) = u(y) exp(i phase), L = ny - 2 rows -- NOT reference behaviour
    (the reference's archived stubs are the phase-0 case, which OracleSolver(bc_y="periodic") restates and the fixtures
    pin); this class defines the phase != 0 semantics the device implements so that the two can be compared: the complex
    field is a real and an imaginary OracleSolver advanced by the reference's real-arithmetic step, coupled only where
    apply_T_pbc / apply_u_pbc copy: a copy from row ny-2 to row 0 multiplies by exp(-i phase), a copy from row 1 to the
    last row by exp(+i phase); every product and sum is a separately rounded float64 operation, real part
    cos*a + sin*b (node rows) / cos*a - sin*b (staggered rows), imaginary part with the sign of sin flipped.
    The source drives the real part only."""

    def __init__(self, x, y, z, C, P, dt, phase, wave="sin", wave_args=None):
        self.re = OracleSolver(x, y, z, C, P, dt, wave=wave, wave_args=wave_args, bc_y="periodic")
        self.im = OracleSolver(x, y, z, C, P, dt, wave=wave, wave_args=wave_args, bc_y="periodic")
        self.c, self.s = float(np.cos(phase)), float(np.sin(phase))
        self.dt = dt

    def _mix(self, a, b, s):
        return self.c * a + s * b

    def step(self):
        re, im, s = self.re, self.im, self.s
        w = SOURCES[re.wave](tt=re.tt, dt=re.dt, **re.wave_args)
        re.uz[0, :, 0] = w
        for o in (re, im):
            o.update_T()
            o.apply_T_tfbc()
        for name in ("T1", "T2", "T3", "T5"):                      # node rows: row 0 <- row ny-2, exp(-i phase)
            a, b = getattr(re, name), getattr(im, name)
            ra, rb = self._mix(a[:, -2, :], b[:, -2, :], s), self._mix(b[:, -2, :], a[:, -2, :], -s)
            a[:, 0, :], b[:, 0, :] = ra, rb
        for name in ("T4", "T6"):                                  # staggered rows: last row <- row 1, exp(+i phase)
            a, b = getattr(re, name), getattr(im, name)
            ra, rb = self._mix(a[:, 1, :], b[:, 1, :], -s), self._mix(b[:, 1, :], a[:, 1, :], s)
            a[:, -1, :], b[:, -1, :] = ra, rb
        for o in (re, im):
            o.update_u()
            o.apply_u_tfbc()
        for name in ("ux_new", "uz_new"):
            a, b = getattr(re, name), getattr(im, name)
            ra, rb = self._mix(a[:, -2, :], b[:, -2, :], s), self._mix(b[:, -2, :], a[:, -2, :], -s)
            a[:, 0, :], b[:, 0, :] = ra, rb
        a, b = re.uy_new, im.uy_new
        ra, rb = self._mix(a[:, 1, :], b[:, 1, :], -s), self._mix(b[:, 1, :], a[:, 1, :], s)
        a[:, -1, :], b[:, -1, :] = ra, rb
        for o in (re, im):
            o.apply_u_abc_xz()
            o.time_step()
            o.tt += 1

    def run(self, steps):
        for _ in range(steps):
            self.step()
        return self
