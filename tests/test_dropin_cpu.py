"""The drop-in claim on the CPU side (no GPU needed):

* files the plugin's writer produces pass a strict structural HDF5 validator written from the format
  specification independently of h5lite (tests/h5check.py), and the validator's own chunk map reads back
  the frames that were written;
* the reference's OWN consumers read those files: simulation/analysis.spectrum (analysis.py:44-96) and the
  read pattern of h5py2gif.py:16-44 / gui/widgets/analysis.py:248-277, through phonomena_b200.h5compat
  installed where they expect h5py (tests marked `ref`: need a copy of the reference);
* plugin discovery: INTEGRATION.md's three-line solver_b200.py dropped into a copy of the reference's
  solvers directory is found by common.findSolvers (common.py:88-110), selected and configured by
  common.loadSettings (common.py:112-156), and Solver.init with the reference's real Grid / Material hands
  the engine the same dt, spacings, inclusion list and Mur coefficients BaseSolver.init works with
  (base_solver.py:194-222); Solver.test() (base_solver.py:286-292) runs on the real TestDefaults objects.
"""
import json
import os
import shutil
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from tests import h5check

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_like_plugin(path, nx=7, ny=6, nz=5, frames=70, mode="surface", seed=3, written=None):
    """A file with the plugin's schema through the plugin's own Writer + the library's native writer threads
    (host-only ring self-test as the frame source: element q of frame f = f * 1e6 + q)."""
    from phonomena_b200.solver_b200 import Writer
    from tests.test_host_cpu import _FakeRingEngine
    rng = np.random.default_rng(seed)
    x = np.cumsum(rng.uniform(0.3, 1.4, nx))
    attrs = {"x": x, "y": np.arange(ny, dtype=float), "z": np.arange(nz, dtype=float),
             "fdx": np.diff(x).reshape(-1, 1, 1), "fdy": np.ones((1, ny - 1, 1)), "fdz": np.ones((1, 1, nz - 1)),
             "sdx": np.ones((nx - 2, 1, 1)), "sdy": np.ones((1, ny - 2, 1)), "sdz": np.ones((1, 1, nz - 2)),
             "steps": frames, "dt": 2.5e-5, "prim_material": "Gallium Arsenide", "sec_material": "Gold",
             "solver_cfg": json.dumps({"record": mode}), "x0": 0, "nxl": nx}
    meta = {"attrs": attrs, "density": rng.uniform(5e3, 2e4, (nx, ny, nz)),
            "elasticity": rng.standard_normal((nx, ny, nz, 6, 6)) if mode == "full" else None}
    eng = _FakeRingEngine(nx, ny, nz, frames if written is None else written)
    w = Writer(path, eng, meta, frames, mode, 1, ring=True)
    w.start()
    w.finish()
    return meta


@pytest.mark.parametrize("mode,frames,written", [("surface", 70, None), ("surface", 200, 131), ("full", 9, None), ("surface", 1, None)])
def test_plugin_files_pass_the_independent_hdf5_validator(tmp_path, mode, frames, written):
    p = str(tmp_path / "v.h5")
    nx, ny, nz = 7, 6, 5
    meta = _write_like_plugin(p, nx, ny, nz, frames, mode, written=written)
    info = h5check.validate(p)
    ds = info["datasets"]
    zext = {"ux": nz, "uy": nz, "uz": nz - 1} if mode == "full" else {"ux": 1, "uy": 1, "uz": 1}
    assert ds["ux"]["shape"] == (nx - 1, ny, zext["ux"], frames) and ds["uy"]["shape"] == (nx, ny - 1, zext["uy"], frames)
    assert ds["uz"]["shape"] == (nx, ny, zext["uz"], frames) and ds["density"]["layout"] == "contiguous"
    got = frames if written is None else written
    assert all(sorted(ds[k]["chunks"]) == list(range(got)) for k in ("ux", "uy", "uz"))      # >64 frames: multi-level B-tree
    a = info["attrs"]
    assert a["steps"] == frames and a["dt"] == 2.5e-5 and a["frames_written"] == got and a["prim_material"] == "Gallium Arsenide"
    assert np.array_equal(a["x"], meta["attrs"]["x"]) and a["fdx"].shape == (nx - 1, 1, 1)
    # frame content through the validator's own chunk map: the ring frame is ux | uy | uz back to back
    n_ux = (nx - 1) * ny * zext["ux"]
    for t in (0, got - 1):
        fr = h5check.read_frame(p, info, "uy", t)
        assert np.array_equal(fr.reshape(-1), t * 1e6 + n_ux + np.arange(fr.size))
    if mode == "full":
        assert ds["elasticity"]["shape"] == (nx, ny, nz, 6, 6)


def test_validator_rejects_damaged_files(tmp_path):
    """The validator is strict: flipping structure bytes of a good file makes it fail (so a pass means something)."""
    p = str(tmp_path / "good.h5")
    _write_like_plugin(p, frames=5)
    raw = bytearray(open(p, "rb").read())
    info = h5check.validate(p)
    root = int.from_bytes(raw[64:72], "little")

    def damaged(mut):
        b = bytearray(raw)
        mut(b)
        q = str(tmp_path / "bad.h5")
        open(q, "wb").write(b)
        with pytest.raises(h5check.H5FormatError):
            h5check.validate(q)

    damaged(lambda b: b.__setitem__(13, 4))                                   # size of offsets
    damaged(lambda b: b.__setitem__(slice(40, 48), (len(raw) + 8).to_bytes(8, "little")))   # EOF address
    damaged(lambda b: b.__setitem__(root, 2))                                 # object header version
    damaged(lambda b: b.__setitem__(root + 16 + 2, b[root + 16 + 2] + 4))     # first message size: misaligned chain
    chunk0 = info["datasets"]["uz"]["chunks"][0]
    damaged(lambda b: b.__setitem__(slice(0, len(b)), b[:chunk0 + 8]))        # truncated file
    bt = raw.find(b"TREE", chunk0)
    damaged(lambda b: b.__setitem__(bt + 4, 7))                               # B-tree node type


def test_h5compat_reads_like_h5py(tmp_path):
    """The h5py-shaped facade: File / get / keys / attrs / Dataset.shape / NumPy indexing incl. the time axis."""
    from phonomena_b200 import h5compat
    p = str(tmp_path / "c.h5")
    nx, ny, frames = 7, 6, 70
    _write_like_plugin(p, nx, ny, 5, frames)
    n_ux, n_uy = (nx - 1) * ny, nx * (ny - 1)
    with h5compat.File(p, "r") as hdf:
        assert sorted(hdf.keys()) == ["density", "ux", "uy", "uz"] and "uz" in hdf and hdf.get("nope") is None
        u = hdf.get("uz")
        assert u != None and u.shape == (nx, ny, 1, frames) and hdf["density"].shape == (nx, ny, 5)      # noqa: E711
        full = np.stack([(t * 1e6 + n_ux + n_uy + np.arange(nx * ny)).reshape(nx, ny, 1) for t in range(frames)], axis=-1)
        assert np.array_equal(u[:, :, 0, :], full[:, :, 0, :]) and np.array_equal(u[:, :, 0, 5], full[:, :, 0, 5])
        assert np.array_equal(u[2, 3, 0, :], full[2, 3, 0, :]) and np.array_equal(u[:, 2, 0, 10:20:3], full[:, 2, 0, 10:20:3])
        assert np.array_equal(u[..., -1], full[..., -1]) and np.array_equal(np.asarray(u), full)
        assert float(np.amin(u[:, :, 0, :])) == float(full.min()) and hdf.attrs["dt"] == 2.5e-5
        assert hdf.attrs["fdx"][:, 0, 0].shape == (nx - 1,)
    with pytest.raises(OSError):
        h5compat.File(p, "w")


@pytest.mark.ref
def test_reference_consumers_read_plugin_output(tmp_path):
    """The reference's own post-processing on a plugin file: simulation/analysis.spectrum (2-D and 1-D, ux and uz) must
    return what the oracle restatement computes from the same arrays, and h5py2gif's read pattern must see the frames."""
    from oracle import refshim, spectrum_numpy
    from phonomena_b200 import h5compat
    refshim.install()
    from simulation import analysis as ranalysis            # the reference module, unmodified
    p = str(tmp_path / "r.h5")
    nx, ny, frames = 9, 6, 64
    meta = _write_like_plugin(p, nx, ny, 5, frames)
    saved = ranalysis.h5py
    ranalysis.h5py = h5compat                                # what `import h5py` would bind on a machine with this facade
    try:
        with h5compat.File(p) as hdf:
            arrays = {k: np.asarray(hdf.get(k)) for k in ("ux", "uz")}
        for u_id in ("ux", "uz"):
            for xi in (None, 3):
                x, f, dft = ranalysis.spectrum(p, u_id, z_index=0, y_index=2, x_index=xi)
                f2, dft2 = spectrum_numpy.spectrum_line(arrays[u_id][:, 2, 0, :], 2.5e-5, x_index=xi)
                x2 = spectrum_numpy.nonlinspace(meta["attrs"]["fdx"][:, 0, 0]) if u_id == "ux" else meta["attrs"]["x"]
                assert np.array_equal(x, x2) and np.array_equal(f, f2) and np.array_equal(dft, dft2), (u_id, xi)
        # an open File object is accepted as well (analysis.py:54)
        with h5compat.File(p, "r") as hdf:
            assert ranalysis.spectrum(hdf, "uz", 0, 2)[2].shape == (nx, frames // 2)
    finally:
        ranalysis.h5py = saved
    # h5py2gif.py:16-24,44 and gui/widgets/analysis.py:250-256
    hdf = h5compat.File(p, "r")
    assert "uz" in list(hdf.keys())
    u, x, y = hdf.get("uz"), hdf.attrs["x"], hdf.attrs["y"]
    size_x, size_y = u.shape[0], u.shape[1]
    X, Y = np.meshgrid(x[:size_x], y[:size_y])
    Z = u[:, :, 0, frames - 1].transpose()
    assert Z.shape == X.shape and np.amax(u[:, :, 0, :]) == Z.max()
    assert hdf.get("density").shape == (nx, ny, 5) and hdf.attrs["z"].shape == (5,) and hdf.attrs["dt"] > 0
    hdf.close()


DISCOVERY = textwrap.dedent('''
    import json, os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np
    from oracle import refshim, fdtd_numpy as onp
    common = refshim.install()                       # PHONOMENA_REF points at the temp copy with solver_b200.py in it

    # a stub engine in place of the ctypes binding: records what Solver.init hands to the device
    from phonomena_b200 import _lib
    calls = {}
    class StubEngine:
        def __init__(self, nx, ny, nz, dt, **kw):
            self.nx, self.ny, self.nz, self.dt, self.kw = nx, ny, nz, dt, kw
            self.x0, self.nxl, self.steps_done, self.launch_count = 0, nx, 0, 0
            calls["create"] = (nx, ny, nz, dt, kw)
        def set_spacing(self, *a): calls["spacing"] = [np.array(v) for v in a]
        def set_material_table(self, c, p): calls["table"] = ([np.array(v) for v in c], list(p))
        def gen_material_ids(self, t, x, y, z): calls["targets"] = (np.array(t), np.array(x), np.array(y), np.array(z))
        def get_material_ids(self):
            t, x, y, z = calls["targets"]
            return onp.material_id_map(x, y, z, onp.make_targets(t.tolist()))
        def set_abc(self, c): calls["abc"] = dict(c) if isinstance(c, dict) else list(c)
        def set_source_table(self, w): calls.setdefault("w", []).extend(np.asarray(w).tolist())
        def run(self, n): self.steps_done += n; self.launch_count += 5 * n
        def sync(self): pass
        def cancel(self): pass
        def info(self): return {"kernel": "stub", "device_bytes": 0}
        def planes(self, c): return self.nx - 1 if c == 0 else self.nx
        def close(self): pass
    _lib.Engine = StubEngine

    common.findSolvers()
    assert "b200" in common.solver_dict and "default" in common.solver_dict, list(common.solver_dict)
    s = common.solver_dict["b200"]
    assert type(s).__module__ == "phonomena_b200.solver_b200" and json.dumps(s.cfg) and s.description.startswith("<p>")

    settings = json.load(open(os.path.join(refshim.REF_ROOT, "data", "default.json")))
    settings["simulation"]["solver"] = "b200"
    settings["simulation"]["cfg"]["write_mode"] = "off"
    settings["simulation"]["cfg"]["precision"] = "fp32"
    path = os.path.join(%(tmp)r, "b200.json")
    json.dump(settings, open(path, "w"))
    cfg, g, m = common.loadSettings(path)
    s = common.solver
    assert s.name == "b200" and s.cfg["precision"] == "fp32" and s.cfg["wave"] == settings["simulation"]["cfg"]["wave"]
    assert s.cfg["record"] == "auto"                 # plugin defaults survive the merge (common.py:149-154)

    # what BaseSolver.init works with (base_solver.py:194-222): the reference solver on the same objects
    r = common.solver_dict["default"]
    r.cfg["write_mode"] = "off"
    r.init(g, m, 7)
    x_before = g.x.copy()
    s.init(g, m, 7)
    assert np.array_equal(g.x, x_before)             # caller's objects untouched
    assert s.dt == r.m.dt, (s.dt, r.m.dt)
    nx, ny, nz, dt, kw = calls["create"]
    assert (nx, ny, nz) == (r.g.x.size, r.g.y.size, r.g.z.size) and dt == r.m.dt and kw["d2"] == r.m.dt ** 2 and kw["dtype"] == "f32"
    for got, ref in zip(calls["spacing"], (r.g.fdx, r.g.fdy, r.g.fdz, r.g.sdx, r.g.sdy, r.g.sdz)):
        assert np.array_equal(np.ravel(got), np.ravel(ref))
    t = calls["targets"][0]
    ref_t = np.array([[q["x"], q["y"], q["z"], q["r"]] for q in r.m.grid.targets], np.float32)
    assert t.dtype == np.float32 and np.array_equal(t, ref_t)
    assert np.array_equal(calls["targets"][1], r.g.x) and np.array_equal(calls["targets"][2], r.g.y)
    tabs, rhos = calls["table"]
    assert np.array_equal(tabs[0], np.array(r.m.primary["c"])) and np.array_equal(tabs[1], np.array(r.m.secondary["c"]))
    assert rhos == [r.m.primary["p"], r.m.secondary["p"]]
    # the id map the device generates from these targets == where the reference put the secondary material
    assert np.array_equal(StubEngine.get_material_ids(None) == 1, r.m.P == r.m.secondary["p"])
    # Mur coefficients: the reference evaluates them inside apply_u_abc (base_solver.py:525-537)
    g_, m_ = r.g, r.m
    vl = np.sqrt(m_.C[0, 0, 0, 0, 0] / m_.P[0, 0, 0]); vt = np.sqrt(m_.C[0, 0, 0, 3, 3] / m_.P[0, 0, 0])
    k = lambda v, d: (v * m_.dt - d) / (v * m_.dt + d)
    want = {"clx": k(vl, g_.sdx[-1, 0, 0]), "ctx": k(vt, g_.fdx[-1, 0, 0]), "cly0": k(vl, g_.sdy[0, 0, 0]), "cty0": k(vt, g_.fdy[0, 0, 0]),
            "cly1": k(vl, g_.sdy[0, -1, 0]), "cty1": k(vt, g_.fdy[0, -1, 0]), "clz": k(vl, g_.sdz[0, 0, -1]), "ctz": k(vt, g_.fdz[0, 0, -1])}
    got = calls["abc"] if isinstance(calls["abc"], dict) else dict(zip(("clx", "ctx", "cly0", "cty0", "cly1", "cty1", "clz", "ctz"), calls["abc"]))
    for key in want:
        assert float(got[key]) == float(want[key]), (key, got[key], want[key])
    # run(): progress / status protocol and the source samples of base_solver.py:251,294-312
    class Sig:
        def __init__(self): self.v = []
        def emit(self, x): self.v.append(x)
    class Signals: pass
    sig = Signals(); sig.status, sig.progress = Sig(), Sig()
    s.run(signals=sig)
    assert sig.progress.v[0] == 0 and sig.progress.v[-1] == 100 and all(type(v) == int for v in sig.progress.v)
    assert sig.status.v[0] == "Solver starting.." and "finished" in sig.status.v[-1]
    wave = {"sin": r.update_sin, "ricker": r.update_ricker}[s.cfg["wave"]]      # the reference's own source functions
    assert calls["w"] == [float(wave(tt=tt, **s.cfg["wave_args"])) for tt in range(7)]

    # Solver.test(): the reference's TestDefaults objects (base_solver.py:28-70,286-292)
    calls.clear()
    s2 = common.importSolver("solver_b200")
    s2.cfg["write_mode"] = "off"
    s2.test()
    from simulation import base_solver
    rt = common.importSolver("solver_default"); rt.cfg["write_mode"] = "off"; rt.init(base_solver.TestDefaults.g, base_solver.TestDefaults.m, 10)
    assert calls["create"][:4] == (rt.g.x.size, rt.g.y.size, rt.g.z.size, rt.m.dt) and s2.stats["steps"] == 10
    print("DISCOVERY OK")
''')


@pytest.mark.ref
def test_plugin_discovery_in_a_reference_checkout(tmp_path):
    from oracle import refshim
    ref = tmp_path / "ref"
    shutil.copytree(os.path.join(refshim.REF_ROOT, "phonomena"), ref / "phonomena", ignore=shutil.ignore_patterns("__pycache__"))
    shutil.copytree(os.path.join(refshim.REF_ROOT, "data"), ref / "data")
    # INTEGRATION.md section 1, verbatim
    (ref / "phonomena" / "simulation" / "solvers" / "solver_b200.py").write_text(
        "# B200 (sm_100a) FDTD solver: same Solver API as solver_default.py, time loop on the GPU\n"
        "from phonomena_b200.solver_b200 import Solver   # noqa: F401\n"
        "from phonomena_b200.solver_b200 import cfg      # noqa: F401  (module-level cfg, like solver_numba.py:8)\n")
    script = tmp_path / "discover.py"
    script.write_text(DISCOVERY % {"root": ROOT, "tmp": str(tmp_path)})
    env = dict(os.environ, PHONOMENA_REF=str(ref), TMPDIR=str(tmp_path))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert r.returncode == 0 and "DISCOVERY OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
