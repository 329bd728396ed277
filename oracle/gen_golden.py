"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference solver
(``solver_default``) -- TEST INFRASTRUCTURE ONLY, runs in the build container
where ``/root/reference`` is mounted:

    python -m oracle.gen_golden            # from the repo root

Each fixture stores the *inputs* a solver needs (mesh lines, float32 inclusion
list, scaled property tables, courant number, source) and the reference's
*outputs* (final ux/uy/uz, T1..T6 of the last step, material id map, dt,
snapshots of the surface uz after selected steps).  The fixtures travel to the
GPU box, the reference does not.
"""
import json
import os
import sys

import numpy as np

from oracle import refshim

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def periodic_y_solver():
    """The reference solver with its ARCHIVED periodic stubs switched on: a subclass of the unmodified
    solver_default.Solver whose update_T_BC / update_u_BC (the two hooks the reference provides for exactly this,
    base_solver.py:374-381, 465-473) call the reference's own apply_T_pbc / apply_u_pbc (:383-400, :475-486), each right
    after the corresponding traction-free update, and then the x and z faces of the Mur ABC (the y faces are what the
    periodic copies replace; their lines of apply_u_abc, :543-550, are left out -- the other six are restated below
    because the reference has them in one function)."""
    refshim.install()
    from simulation.solvers import solver_default

    class PeriodicY(solver_default.Solver):
        def update_T_BC(self):
            self.apply_T_tfbc()
            self.apply_T_pbc()

        def update_u_BC(self):
            self.apply_u_tfbc()
            self.apply_u_pbc()
            g, m = self.g, self.m
            vl = np.sqrt(m.C[0, 0, 0, 0, 0] / m.P[0, 0, 0])
            vt = np.sqrt(m.C[0, 0, 0, 3, 3] / m.P[0, 0, 0])
            ctx = ((vt * m.dt - g.fdx) / (vt * m.dt + g.fdx))[:, 0, 0]
            clx = ((vl * m.dt - g.sdx) / (vl * m.dt + g.sdx))[:, 0, 0]
            ctz = ((vt * m.dt - g.fdz) / (vt * m.dt + g.fdz))[0, 0, :]
            clz = ((vl * m.dt - g.sdz) / (vl * m.dt + g.sdz))[0, 0, :]
            g.ux_new[-1, :, :] = g.ux[-2, :, :] + clx[-1] * (g.ux_new[-2, :, :] - g.ux[-1, :, :])
            g.uy_new[-1, :, :] = g.uy[-2, :, :] + ctx[-1] * (g.uy_new[-2, :, :] - g.uy[-1, :, :])
            g.uz_new[-1, :, :] = g.uz[-2, :, :] + ctx[-1] * (g.uz_new[-2, :, :] - g.uz[-1, :, :])
            g.ux_new[:, :, -1] = g.ux[:, :, -2] + ctz[-1] * (g.ux_new[:, :, -2] - g.ux[:, :, -1])
            g.uy_new[:, :, -1] = g.uy[:, :, -2] + ctz[-1] * (g.uy_new[:, :, -2] - g.uy[:, :, -1])
            g.uz_new[:, :, -1] = g.uz[:, :, -2] + clz[-1] * (g.uz_new[:, :, -2] - g.uz[:, :, -1])

    s = PeriodicY()
    s.cfg["write_mode"] = "off"
    return s


def _run(grid, material, cfg, steps, snaps=(), solver=None):
    s = solver if solver is not None else refshim.default_solver()
    s.cfg.update(cfg)
    s.cfg["write_mode"] = "off"
    s.init(grid, material, steps)
    snaps = sorted(set(snaps))
    frames = {}
    # drive the reference loop one step at a time through its own methods, in its own order
    # (base_solver.py:245-260) so that intermediate surface frames can be captured.
    wave_fn = {"sin": s.update_sin, "ricker": s.update_ricker}[s.cfg["wave"]]
    for tt in range(steps):
        s.g.uz[0, :, 0] = wave_fn(tt=tt, **s.cfg["wave_args"])
        s.update_T()
        s.update_T_BC()
        s.update_u()
        s.update_u_BC()
        s.time_step()
        if (tt + 1) in snaps:
            frames[tt + 1] = (s.g.ux[:, :, 0].copy(), s.g.uy[:, :, 0].copy(), s.g.uz[:, :, 0].copy())
    return s, frames


def _pack(name, s, frames, steps, courant, extra=None):
    g, m = s.g, s.m
    # primary = value at a cell that is never an inclusion; ids from P/C comparison
    prim_c, sec_c = np.array(m.primary["c"], float), np.array(m.secondary["c"], float)
    prim_p, sec_p = float(m.primary["p"]), float(m.secondary["p"])
    if prim_p != sec_p:
        ids = (m.P == sec_p).astype(np.uint8)
    else:
        ids = np.zeros(m.P.shape, np.uint8)
    assert np.array_equal(np.where(ids[..., None, None] == 1, sec_c, prim_c), m.C)
    targets = np.array([[t["x"], t["y"], t["z"], t["r"]] for t in m.grid.targets], np.float32).reshape(-1, 4)
    d = dict(
        x=g.x, y=g.y, z=g.z, targets=targets,
        prim_c=prim_c, prim_p=prim_p, sec_c=sec_c, sec_p=sec_p,
        courant=float(courant), dt=float(m.dt), steps=int(steps),
        wave=str(s.cfg["wave"]), wave_args=json.dumps(s.cfg["wave_args"]),
        ids=ids, ux=g.ux, uy=g.uy, uz=g.uz,
        T1=g.T1, T2=g.T2, T3=g.T3, T4=g.T4, T5=g.T5, T6=g.T6,
        ux_old=g.ux_old, uy_old=g.uy_old, uz_old=g.uz_old,
        snap_steps=np.array(sorted(frames), np.int64),
    )
    for k, (sx, sy, sz) in frames.items():
        d["snap_ux_%d" % k], d["snap_uy_%d" % k], d["snap_uz_%d" % k] = sx, sy, sz
    if extra:
        d.update(extra)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **d)
    print("%-22s grid %s steps %d dt %.17g |uz| %.17g  -> %s (%.0f KB)" % (
        name, m.P.shape, steps, m.dt, np.linalg.norm(g.uz), os.path.relpath(path), os.path.getsize(path) / 1024))


def case_testdefaults():
    """Solver.test(): TestDefaults grid, 10 steps, ricker (base_solver.py:28-70,286-292)."""
    refshim.install()
    from simulation import base_solver
    s, fr = _run(base_solver.TestDefaults.g, base_solver.TestDefaults.m, {"wave": "ricker", "wave_args": {"f": 100}}, 10, snaps=(1, 10))
    _pack("testdefaults", s, fr, 10, s.m.c_max)


def case_settings(name, path, steps, snaps, cfg_over=None):
    common = refshim.install()
    common.findSolvers()
    cfg, g, m = common.loadSettings(path)
    scfg = dict(cfg["simulation"]["cfg"])
    scfg.update(cfg_over or {})
    s, fr = _run(g, m, scfg, steps, snaps)
    _pack(name, s, fr, steps, cfg["simulation"]["courant"])


def case_custom(name, size, max_d, min_d, incl, steps, snaps, wave, wave_args, courant=0.1,
                primary="GaAs", secondary="Au", periodic_y=False):
    common = refshim.install()
    from simulation import grid as rgrid, material as rmat
    props = json.load(open(os.path.join(refshim.REF_ROOT, "data", "default.json")))["material"]["properties"]
    g = rgrid.Grid()
    g.init(size_x=size[0], size_y=size[1], size_z=size[2])
    g.min_d = min_d
    g.max_dx, g.max_dy, g.max_dz = max_d
    g.slope = 1.0
    for (x, y, z, r) in incl:
        g.addInclusion(x=x, y=y, z=z, r=r)
    g.buildMesh()
    g.update()
    m = rmat.Material()
    m.init(grid=g, properties=props)
    m.c_max = courant
    m.setPrimary(primary)
    m.setSecondary(secondary)
    m.update()
    s, fr = _run(g, m, {"wave": wave, "wave_args": wave_args}, steps, snaps, solver=periodic_y_solver() if periodic_y else None)
    _pack(name, s, fr, steps, courant, extra={"bc_y": "periodic"} if periodic_y else None)


def case_spectrum(name, path, steps, y_index, x_index):
    """The reference's own post-processing (simulation/analysis.py:44-96) on the reference solver's
    own frames: run `steps` steps, keep the (x, t) lines a probe would keep (frame tt = the state
    after time_step of step tt, base_solver.py:256-260), hand them to `spectrum` through the
    in-memory h5py stand-in and store lines + answers."""
    common = refshim.install()
    common.findSolvers()
    cfg, g, m = common.loadSettings(path)
    s = refshim.default_solver()
    s.cfg.update(dict(cfg["simulation"]["cfg"]))
    s.cfg["write_mode"] = "off"
    s.init(g, m, steps)
    wave_fn = {"sin": s.update_sin, "ricker": s.update_ricker}[s.cfg["wave"]]
    full = {k: np.zeros(getattr(s.g, k).shape + (steps,)) for k in ("ux", "uy", "uz")}
    for tt in range(steps):
        s.g.uz[0, :, 0] = wave_fn(tt=tt, **s.cfg["wave_args"])
        s.update_T(); s.update_T_BC(); s.update_u(); s.update_u_BC(); s.time_step()
        for k in full:
            full[k][..., tt] = getattr(s.g, k)
    import h5py                                   # the stand-in installed by refshim
    from simulation import analysis as ranalysis
    f = h5py.File.in_memory(full, {"x": s.g.x, "fdx": s.g.fdx, "dt": s.m.dt})
    d = dict(steps=steps, dt=s.m.dt, y_index=y_index, x_index=x_index, x=s.g.x, fdx=s.g.fdx[:, 0, 0])
    for u_id in ("ux", "uz"):
        d["line_" + u_id] = full[u_id][:, y_index, 0, :]
        for tag, xi in (("2d", None), ("1d", x_index)):
            x, fr, dft = ranalysis.spectrum(f, u_id, z_index=0, y_index=y_index, x_index=xi)
            d["x_%s" % u_id], d["f"], d["dft_%s_%s" % (u_id, tag)] = x, fr, dft
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print("wrote", name, {k: np.shape(v) for k, v in d.items()})


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = refshim.REF_ROOT
    case_testdefaults()
    case_settings("default_json_1000", os.path.join(ref, "data", "default.json"), 1000, (1, 2, 3, 10, 100, 1000))
    case_settings("nonuniform_json_200", os.path.join(ref, "tests", "data", "nonuniform.json"), 200, (1, 50, 200))
    case_settings("fine_json_200", os.path.join(ref, "tests", "data", "fine.json"), 200, (200,))
    # homogeneous GaAs cube, the shape of BASELINE config #2 at reduced size
    case_custom("homog_32_20", (31, 31, 31), (1, 1, 1), 1, [], 20, (5, 20), "sin", {"f": 100}, secondary="GaAs")
    # partial-depth inclusions with dz = 0.5: exercises the z re-binding quirk of
    # inclusionIndices (grid.py:173) and a delayed ricker source
    case_custom("partial_depth_dz05", (24, 20, 6), (1, 1, 0.5), 0.5,
                [(8.0, 10.0, 3.0, 2.5), (17.0, 9.0, 2.0, 2.0)], 60, (1, 60), "ricker",
                {"f": 2000, "source_delay": 2e-4}, secondary="Al")
    # small crystal, the shape of BASELINE config #3 at reduced size (uniform integer mesh)
    case_custom("crystal_48x32x12", (47, 31, 11), (1, 1, 1), 1,
                [(8.0 + 16 * a, 8.0 + 16 * b, 11.0, 4.0) for a in range(3) for b in range(2)][:5],
                80, (1, 80), "sin", {"f": 100})
    # periodic y boundaries: the reference's archived stubs switched on (periodic_y_solver above);
    # a crystal strip one lattice period wide in y with a non-uniform x mesh, and a homogeneous block
    case_custom("periodic_y_crystal_40x18x14", (39, 17, 13), (1.4, 1, 1), 0.5,
                [(10.0, 8.0, 13.0, 3.5), (26.0, 8.0, 6.0, 3.0)], 120, (1, 2, 40, 120), "sin", {"f": 100}, periodic_y=True)
    case_custom("periodic_y_homog_24x12x10", (23, 11, 9), (1, 1, 1), 1, [], 60, (1, 60), "ricker", {"f": 100},
                secondary="GaAs", periodic_y=True)
    case_spectrum("spectrum_default_json_128", os.path.join(ref, "data", "default.json"), 128, 10, 5)


if __name__ == "__main__":
    sys.exit(main())
