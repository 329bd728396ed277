"""The plugin (reference interface) end to end on the GPU: Solver.init/run/cancel, signals,
surface and full-field HDF5 output, compared with the reference's golden outputs."""
import json
import threading
import time
from types import SimpleNamespace

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


def fake_from_golden(d):
    """Grid / Material stand-ins with the attributes Solver.init reads, from a fixture's inputs."""
    tdt = np.dtype([("x", "f"), ("y", "f"), ("z", "f"), ("r", "f")])
    t = np.array([tuple(r) for r in d["targets"]], dtype=tdt).reshape(-1)
    g = SimpleNamespace(x=d["x"].copy(), y=d["y"].copy(), z=d["z"].copy(), targets=t, SI_conversion=1)
    m = SimpleNamespace(primary={"c": d["prim_c"], "p": d["prim_p"], "name": "P"},
                        secondary={"c": d["sec_c"], "p": d["sec_p"], "name": "S"}, c_max=d["courant"], grid=g)
    return g, m


class Signals:
    def __init__(self):
        self.p, self.s = [], []
        self.progress = SimpleNamespace(emit=self.p.append)
        self.status = SimpleNamespace(emit=self.s.append)


def make_solver(d, tmp_path, **cfg):
    from phonomena_b200.solver_b200 import Solver
    s = Solver()
    s.cfg.update({"wave": d["wave"], "wave_args": d["wave_args"], "write_mode": "thread", "arith": "exact"})
    s.cfg.update(cfg)
    s.file = str(tmp_path / "out.h5")
    return s


@pytest.mark.parametrize("name", ["default_json_1000", "nonuniform_json_200", "partial_depth_dz05"])
def test_plugin_surface_recording_matches_reference(name, tmp_path):
    from phonomena_b200.h5lite import H5Reader
    d = H.load_golden(name)
    s = make_solver(d, tmp_path, record="surface", chunk_steps=37)
    g, m = fake_from_golden(d)
    sig = Signals()
    s.init(g, m, d["steps"])
    assert s.dt == d["dt"]
    s.run(signals=sig)
    assert sig.p[0] == 0 and sig.p[-1] == 100 and all(isinstance(v, int) for v in sig.p)
    assert sig.s[0] == "Solver starting.." and "finished" in sig.s[-1]
    for got, key in zip(s.fields(), ("ux", "uy", "uz")):
        assert np.array_equal(got, d[key]), key
    r = H5Reader(s.file)
    nx, ny, nz = d["ids"].shape
    assert r.shape("uz") == (nx, ny, 1, d["steps"]) and r.shape("ux") == (nx - 1, ny, 1, d["steps"])
    assert r.attrs["steps"] == d["steps"] and r.attrs["dt"] == d["dt"] and r.attrs["frames_written"] == d["steps"]
    assert np.array_equal(r.attrs["x"], d["x"]) and r.attrs["fdx"].shape == (nx - 1, 1, 1)
    assert json.loads(r.attrs["solver_cfg"])["record"] == "surface"
    P = np.where(d["ids"] == 1, d["sec_p"], d["prim_p"])
    assert np.array_equal(r.read("density"), P)
    # frame t of the file = state after step t+1 (App. B #10), z-index 0
    for n in (int(v) for v in d["snap_steps"]):
        assert np.array_equal(r.read("uz", frame=n - 1)[:, :, 0], d["snap_uz_%d" % n]), n
        assert np.array_equal(r.read("ux", frame=n - 1)[:, :, 0], d["snap_ux_%d" % n]), n
        assert np.array_equal(r.read("uy", frame=n - 1)[:, :, 0], d["snap_uy_%d" % n]), n


def test_plugin_full_field_output_matches_reference_schema(tmp_path):
    from phonomena_b200.h5lite import H5Reader
    d = H.load_golden("testdefaults")
    s = make_solver(d, tmp_path, record="full")
    g, m = fake_from_golden(d)
    s.init(g, m, d["steps"])
    s.run()
    r = H5Reader(s.file)
    nx, ny, nz = d["ids"].shape
    assert r.shape("ux") == (nx - 1, ny, nz, 10) and r.shape("uy") == (nx, ny - 1, nz, 10) and r.shape("uz") == (nx, ny, nz - 1, 10)
    assert sorted(r.datasets) == ["density", "elasticity", "ux", "uy", "uz"]
    assert r.read("elasticity").shape == (nx, ny, nz, 6, 6)
    for key in ("ux", "uy", "uz"):
        assert np.array_equal(r.read(key, frame=9), d[key]), key


@pytest.mark.parametrize("every", [1, 3])
def test_plugin_full_field_ring_and_synchronous_paths_agree(tmp_path, every):
    """record = "full": frames stream through the recorder ring when they are small (the default for the reference's
    grid sizes) and are read back synchronously otherwise -- same file either way, and the k = 0 plane of frame t is
    the reference's surface snapshot of step t + 1."""
    from phonomena_b200.h5lite import H5Reader
    d = H.load_golden("crystal_48x32x12")
    files = []
    for ring in (True, False):
        s = make_solver(d, tmp_path, record="full", record_every=every, chunk_steps=7)
        s.file = str(tmp_path / ("full_%d.h5" % ring))
        if not ring:
            s.FULL_RING_MAX_FRAME_BYTES = 0
        g, m = fake_from_golden(d)
        s.init(g, m, d["steps"])
        assert s._full_ring == ring
        s.run()
        files.append(s.file)
    a, b = H5Reader(files[0]), H5Reader(files[1])
    frames = d["steps"] // every
    for key in ("ux", "uy", "uz"):
        assert a.shape(key) == b.shape(key) and a.shape(key)[-1] == frames
        for t in range(frames):
            assert np.array_equal(a.read(key, frame=t), b.read(key, frame=t)), (key, t)
    assert a.attrs["frames_written"] == b.attrs["frames_written"] == frames
    for n in (int(v) for v in d["snap_steps"]):
        if n % every == 0:
            assert np.array_equal(a.read("uz", frame=n // every - 1)[:, :, 0], d["snap_uz_%d" % n]), n
    if d["steps"] % every == 0:
        for key in ("ux", "uy", "uz"):
            assert np.array_equal(a.read(key, frame=frames - 1), d[key]), key


def test_plugin_write_mode_off_and_fp32(tmp_path):
    d = H.load_golden("crystal_48x32x12")
    s = make_solver(d, tmp_path, write_mode="off", precision="fp32", arith="fast")
    g, m = fake_from_golden(d)
    s.init(g, m, d["steps"])
    s.run()
    assert H.rel_l2(s.fields(), [d["ux"], d["uy"], d["uz"]]) <= 1e-5
    assert s.stats["steps"] == d["steps"] and s.stats["launches"] > 0


def test_plugin_cancel_from_another_thread(tmp_path):
    d = H.load_golden("crystal_48x32x12")
    s = make_solver(d, tmp_path, write_mode="off", chunk_steps=5)
    g, m = fake_from_golden(d)
    s.init(g, m, 2_000_000)
    th = threading.Thread(target=s.run)
    th.start()
    time.sleep(0.3)
    s.cancel()
    th.join(5)
    assert not th.is_alive() and 0 < s.stats["steps"] < 2_000_000


def test_plugin_does_not_mutate_caller_objects(tmp_path):
    d = H.load_golden("default_json_1000")
    s = make_solver(d, tmp_path, write_mode="off")
    g, m = fake_from_golden(d)
    x0 = g.x.copy()
    s.init(g, m, 3)
    g.x[:] = -1            # the caller may free / reuse its arrays after init()
    s.run()
    assert np.array_equal(x0, d["x"]) and s.stats["steps"] == 3


def _arrays_material(d, C, P):
    g, m = fake_from_golden(d)
    m.C, m.P = C, P
    return g, m


def test_plugin_material_from_reference_arrays(tmp_path):
    """cfg material='arrays': Material.C / Material.P as the reference stores them."""
    from oracle import fdtd_numpy as onp
    d = H.load_golden("default_json_1000")
    C, P = onp.set_constants(d["x"], d["y"], d["z"], H.targets_of(d), d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    s = make_solver(d, tmp_path, record="full", material="arrays")
    g, m = _arrays_material(d, C, P)
    g.targets = g.targets[:0]                  # the inclusion list must not be what is used
    s.init(g, m, 200)
    s.run()
    from phonomena_b200.h5lite import H5Reader
    r = H5Reader(s.file)
    assert np.array_equal(r.read("density"), P) and np.array_equal(r.read("elasticity"), C)
    assert np.array_equal(r.read("uz", frame=99)[:, :, 0], d["snap_uz_100"])
    assert np.array_equal(r.read("ux", frame=9)[:, :, 0], d["snap_ux_10"])


@pytest.mark.parametrize("precision,arith,tol", [("fp64", "exact", 0.0), ("fp64", "fast", 1e-12), ("fp32", "fast", 1e-5)])
def test_plugin_three_material_medium_matches_oracle(tmp_path, precision, arith, tol):
    """More than the reference's two materials, straight from C / P: a third (and fourth) material in
    blocks that cut through inclusions, the corner cell and the boundary planes."""
    d = H.load_golden("crystal_48x32x12")
    o = H.oracle_from_golden(d)
    C, P = o.C.copy(), o.P.copy()
    third = d["sec_c"] * 0.37
    third[0, 2] *= 1.11                        # keep it non-symmetric like the shipped tables (App. B #4)
    C[5:19, 3:17, :7], P[5:19, 3:17, :7] = third, 2700.0
    C[0:3, 0:4, 0:2], P[0:3, 0:4, 0:2] = d["prim_c"] * 1.5, 4000.0      # the Mur corner cell
    C[-2:, :, -3:], P[-2:, :, -3:] = third, 2700.0
    from oracle import fdtd_numpy as onp
    o2 = onp.OracleSolver(d["x"], d["y"], d["z"], C, P, o.dt, wave=d["wave"], wave_args=d["wave_args"])
    o2.run(60)
    s = make_solver(d, tmp_path, record="off", write_mode="off", material="arrays", precision=precision, arith=arith)
    g, m = _arrays_material(d, C, P)
    s.init(g, m, 60)
    assert s.dt == o.dt
    s.run()
    got, ref = s.fields(), (o2.ux, o2.uy, o2.uz)
    if tol == 0.0:
        assert all(np.array_equal(a, b) for a, b in zip(got, ref))
    else:
        assert H.rel_l2(got, ref) <= tol


def test_plugin_material_arrays_errors(tmp_path):
    d = H.load_golden("testdefaults")
    o = H.oracle_from_golden(d)
    s = make_solver(d, tmp_path, record="off", write_mode="off", material="arrays")
    g, m = _arrays_material(d, o.C[:-1], o.P[:-1])
    with pytest.raises(ValueError, match="shapes"):
        s.init(g, m, 2)
    C, P = o.C.copy(), o.P.copy()
    P.reshape(-1)[:40] += np.arange(40)        # 40 distinct densities
    g, m = _arrays_material(d, C, P)
    with pytest.raises(Exception, match="distinct materials"):
        s.init(g, m, 2)


def test_plugin_record_fields_subset(tmp_path):
    """cfg["record_fields"]: only the selected components are recorded (BASELINE config #3 records surface u_z)."""
    from phonomena_b200.h5lite import H5Reader
    d = H.load_golden("default_json_1000")
    for mode in ("surface", "full"):
        s = make_solver(d, tmp_path, record=mode, record_fields=["uz"], record_every=100)
        s.file = str(tmp_path / ("uz_%s.h5" % mode))
        g, m = fake_from_golden(d)
        s.init(g, m, d["steps"])
        s.run()
        r = H5Reader(s.file)
        assert "uz" in r.datasets and "ux" not in r.datasets and "uy" not in r.datasets
        assert r.shape("uz")[-1] == d["steps"] // 100
        last = r.read("uz", frame=d["steps"] // 100 - 1)
        assert np.array_equal(last[:, :, 0], d["uz"][:, :, 0])
        if mode == "full":
            assert np.array_equal(last, d["uz"])
    with pytest.raises(ValueError):
        s = make_solver(d, tmp_path, record_fields=[])
        s.init(*fake_from_golden(d), 10)


# ---- the recorder cannot hang run() (reference: queue.get(timeout=120), join(300), base_solver.py:89-92,148,274) ----
def test_plugin_writer_failure_raises_instead_of_hanging(tmp_path, monkeypatch):
    """The writer threads hit a write error on the first frame (their descriptor is swapped for a read-only one):
    Solver.run() must raise with the errno text within about a second, not wait for a free ring slot forever."""
    import os
    from phonomena_b200 import _lib
    from phonomena_b200.solver_b200 import Writer
    d = H.load_golden("crystal_48x32x12")
    s = make_solver(d, tmp_path, record="surface", chunk_steps=20)
    monkeypatch.setattr(Writer, "PREALLOCATE_MAX_BYTES", 0)      # unallocated extents -> the threads pwrite() (allocated ones are mapped)
    s.init(*fake_from_golden(d), 400)            # 400 frames >> 32 ring slots
    assert not s.engine.writer_mapped()
    ro = os.open(os.devnull, os.O_RDONLY)
    keep = os.dup(s.writer.h5.fd)
    os.dup2(ro, s.writer.h5.fd)                  # native pwrite -> EBADF
    t0 = time.time()
    with pytest.raises(_lib.PhbError, match="pwrite"):
        s.run()
    assert time.time() - t0 < 3.0
    os.dup2(keep, s.writer.h5.fd)
    os.close(keep)
    os.close(ro)
    with pytest.raises(RuntimeError, match="init"):
        s.run()                                  # a second run() without init() is refused, not hung


def test_recorder_full_ring_times_out_and_cancel_breaks_in():
    """No consumer at all: phb_run fails after the configured wait; phb_cancel from another thread ends a wait in progress."""
    from phonomena_b200 import _lib
    d = H.load_golden("crystal_48x32x12")
    with H.engine_from_golden(d, record_mask=_lib.REC_UZ, ring_slots=4) as e:
        e.record_timeout(300)
        t0 = time.time()
        with pytest.raises(_lib.PhbError, match="ring full"):
            e.run(50)
        assert 0.25 < time.time() - t0 < 2.0 and e.steps_done == 5      # 4 frames fit, the 5th step's frame does not
    with H.engine_from_golden(d, record_mask=_lib.REC_UZ, ring_slots=4) as e:
        e.record_timeout(60000)
        threading.Timer(0.3, e.cancel).start()
        t0 = time.time()
        with pytest.raises(_lib.PhbCancelled):
            e.run(50)
        assert time.time() - t0 < 2.0
        # the consumer API still drains what was recorded
        tt, views = e.record_next(timeout_ms=1000)
        assert tt == 0 and views["uz"].shape == (48, 32)
        e.record_release()
        e.record_abort("test")
        with pytest.raises(_lib.PhbError, match="aborted"):
            e.run(1)


def test_plugin_reinit_without_run_and_cancel_latency(tmp_path):
    """init() twice with no run() in between: the first writer is stopped and its file closed before the engine it
    reads from is destroyed (was a use-after-free).  cancel() stops a long chunk within a step, not a chunk."""
    from phonomena_b200.h5lite import H5Reader
    d = H.load_golden("crystal_48x32x12")
    s = make_solver(d, tmp_path, record="surface")
    s.file = str(tmp_path / "first.h5")
    s.init(*fake_from_golden(d), 30)
    assert s.engine.writer_mapped()              # allocated extents: the native threads copy through a shared mapping
    first = s.file
    s.file = str(tmp_path / "second.h5")
    s.init(*fake_from_golden(d), 30)
    assert H5Reader(first).attrs["frames_written"] == 0
    s.run()
    assert H5Reader(s.file).attrs["frames_written"] == 30
    s2 = make_solver(d, tmp_path, write_mode="off", chunk_steps=100_000)       # (the host evaluates a chunk's source samples first: 2 us each)
    s2.init(*fake_from_golden(d), 10_000_000)
    th = threading.Thread(target=s2.run)
    th.start()
    time.sleep(0.3)
    t0 = time.time()
    s2.cancel()
    th.join(5)
    assert not th.is_alive() and time.time() - t0 < 1.0 and 0 < s2.stats["steps"] < 10_000_000


@pytest.mark.parametrize("ring", [True, False])
def test_plugin_decimated_volume_snapshots(tmp_path, ring):
    """record = "full" with cfg["record_stride"]: every (sx, sy, sz)-th entry of the whole arrays per recorded step --
    the way to keep volume snapshots of a grid whose full frames would not fit the recorder ring.
    The file describes the decimated mesh (x, fdx, ... match the dataset shapes, the full mesh lines are kept as
    x_full ...); the frames equal the reference's fields on the kept entries; ring and synchronous paths agree."""
    from phonomena_b200.h5lite import H5Reader
    d = H.load_golden("crystal_48x32x12")
    st = (2, 3, 2)
    s = make_solver(d, tmp_path, record="full", record_stride=list(st), record_every=10)
    if not ring:
        s.FULL_RING_MAX_FRAME_BYTES = 0
    s.init(*fake_from_golden(d), d["steps"])
    assert s._full_ring == ring
    s.run()
    r = H5Reader(s.file)
    frames = d["steps"] // 10
    for key in ("ux", "uy", "uz"):
        want = d[key][::st[0], ::st[1], ::st[2]]
        assert r.shape(key) == want.shape + (frames,), (key, r.shape(key))
        assert np.array_equal(r.read(key, frame=frames - 1), want), key
    assert np.array_equal(r.attrs["x"], d["x"][::2]) and np.array_equal(r.attrs["x_full"], d["x"])
    assert r.attrs["fdx"].shape == (len(d["x"][::2]) - 1, 1, 1) and list(r.attrs["record_stride"]) == [2.0, 3.0, 2.0]
    P = np.where(d["ids"] == 1, d["sec_p"], d["prim_p"])
    assert np.array_equal(r.read("density"), P[::2, ::3, ::2])
    from tests import h5check
    h5check.validate(s.file)


def test_plugin_default_record_mode_is_the_reference_schema_on_small_grids(tmp_path):
    """cfg["record"] = "auto" (the default): whole fields per step, i.e. exactly what the reference's Writer stores
    (base_solver.py:105-133, 135-160), while a frame is small; the surface planes only above the size limit."""
    from phonomena_b200.h5lite import H5Reader
    d = H.load_golden("testdefaults")
    nx, ny, nz = d["ids"].shape
    s = make_solver(d, tmp_path)
    assert s.cfg["record"] == "auto"
    s.init(*fake_from_golden(d), d["steps"])
    s.run()
    r = H5Reader(s.file)
    assert r.attrs["record"] == "full" and r.shape("uz") == (nx, ny, nz - 1, d["steps"]) and "elasticity" in r.datasets
    assert np.array_equal(r.read("ux", frame=d["steps"] - 1), d["ux"])
    s2 = make_solver(d, tmp_path)
    s2.AUTO_FULL_MAX_FRAME_BYTES = 0
    s2.file = str(tmp_path / "big.h5")
    s2.init(*fake_from_golden(d), d["steps"])
    s2.run()
    r2 = H5Reader(s2.file)
    assert r2.attrs["record"] == "surface" and r2.shape("uz") == (nx, ny, 1, d["steps"])
    with pytest.raises(ValueError, match="record"):
        make_solver(d, tmp_path, record="everything").init(*fake_from_golden(d), 3)
