"""CPU-only checks: host math equals the oracle bit for bit, slab split, HDF5 writer/reader,
the C-ABI library loads and exports every symbol include/phb200.h declares, and the product
fails loudly without a GPU (no CPU fallback)."""
import os
import re

import numpy as np
import pytest

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", H.golden_names())
def test_hostmath_matches_oracle(name):
    from oracle import fdtd_numpy as onp
    from phonomena_b200 import hostmath as hm
    d = H.load_golden(name)
    o = H.oracle_from_golden(d)
    sp = hm.spacings(d["x"], d["y"], d["z"])
    for a, b in zip(sp, (o.fdx, o.fdy, o.fdz, o.sdx, o.sdy, o.sdz)):
        assert np.array_equal(a, b.reshape(-1))
    dt = hm.cfl_dt(sp[0], sp[1], sp[2], d["courant"], {"c": d["prim_c"], "p": d["prim_p"]}, {"c": d["sec_c"], "p": d["sec_p"]})
    assert dt == d["dt"] == o.dt
    corner = (d["sec_c"], d["sec_p"]) if d["ids"][0, 0, 0] else (d["prim_c"], d["prim_p"])
    assert hm.abc_coefficients(corner[0], corner[1], dt, *sp) == {k: float(v) for k, v in o.abc_coefficients().items()}
    w = hm.source_table(d["wave"], 20, dt, d["wave_args"])
    assert np.array_equal(w, onp.source_table(d["wave"], 20, dt, d["wave_args"]))
    assert np.array_equal(hm.source_table(d["wave"], 5, dt, d["wave_args"], start=15), w[15:])


def test_split_slabs():
    from phonomena_b200 import hostmath as hm
    assert hm.split_slabs(4096, 8) == [(512 * r, 512) for r in range(8)]
    s = hm.split_slabs(1030, 4)
    assert sum(n for _, n in s) == 1030 and s[0][0] == 0 and all(a[0] + a[1] == b[0] for a, b in zip(s, s[1:]))
    with pytest.raises(ValueError):
        hm.split_slabs(10, 4)


def test_library_exports_every_declared_symbol():
    from phonomena_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "phb200.h")).read()
    declared = set(re.findall(r"\b(phb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = _lib.load_library()
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.phb_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from phonomena_b200 import _lib
    with pytest.raises(_lib.PhbError, match="no CPU fallback"):
        _lib.Engine(8, 8, 8, 1e-5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "phonomena_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|#include.*oracle|dlopen.*oracle|CDLL.*oracle", src, re.M), f


def test_h5lite_roundtrip_multilevel_btree(tmp_path):
    from phonomena_b200.h5lite import H5Reader, H5Writer, SIG
    rng = np.random.default_rng(1)
    p = str(tmp_path / "t.h5")
    frames = [rng.standard_normal((7, 5, 1)) for _ in range(5000)]      # > 64*64 chunks -> 3-level B-tree
    with H5Writer(p) as w:
        w.attrs.update({"x": np.arange(7.0), "fdx": np.ones((6, 1, 1)), "steps": 5000, "dt": 2.5e-5, "prim_material": "Gallium Arsenide"})
        P = rng.standard_normal((7, 5, 3))
        w.create_dataset("density", P)
        d = w.create_chunked("uz", (7, 5, 1, 5000))
        for t, f in enumerate(frames):
            w.write_frame(d, t, f)
    raw = open(p, "rb").read()
    assert raw[:8] == SIG and int.from_bytes(raw[40:48], "little") == len(raw)      # signature, EOF address
    r = H5Reader(p)
    assert r.attrs["steps"] == 5000 and r.attrs["dt"] == 2.5e-5 and r.attrs["prim_material"] == "Gallium Arsenide"
    assert r.attrs["fdx"].shape == (6, 1, 1) and r.shape("uz") == (7, 5, 1, 5000)
    assert np.array_equal(r.read("density"), P)
    for t in (0, 63, 64, 4095, 4096, 4999):
        assert np.array_equal(r.read("uz", frame=t), frames[t])


def test_spectrum_matches_reference_formula(tmp_path):
    """phonomena_b200.analysis.spectrum on an h5lite file == the reference's expression
    (simulation/analysis.py:59-88) evaluated with NumPy on the same data."""
    from phonomena_b200 import analysis
    from phonomena_b200.h5lite import H5Writer
    rng = np.random.default_rng(2)
    nx, ny, N, dt = 9, 5, 64, 2.5e-5
    data = rng.standard_normal((nx, ny, 1, N))
    x = np.cumsum(rng.uniform(0.5, 1.5, nx))
    p = str(tmp_path / "s.h5")
    with H5Writer(p) as w:
        w.attrs.update({"x": x, "fdx": np.diff(x).reshape(-1, 1, 1), "dt": dt, "steps": N})
        d = w.create_chunked("uz", (nx, ny, 1, N))
        for t in range(N):
            w.write_frame(d, t, data[..., t])
    xs, f, dft = analysis.spectrum(p, "uz", 0, 3)
    win = np.hanning(N)
    ref = np.abs(np.fft.fft2(data[:, 3, 0, :] * win, norm="ortho"))[:, :N // 2]
    assert np.array_equal(xs, x) and np.array_equal(f, np.fft.fftfreq(N, d=dt)[:N // 2]) and np.allclose(dft, ref, rtol=1e-13, atol=0)
    _, _, d1 = analysis.spectrum(p, "uz", 0, 3, x_index=4)
    assert np.allclose(d1, np.abs(np.fft.fft(data[4, 3, 0, :] * win, norm="ortho"))[:N // 2], rtol=1e-13, atol=0)


def test_bench_reference_arm_contract_line():
    """`bench.py --impl reference` (CPU only): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                        "--quick-reference"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "Gcell-updates/s" and line["unit"] == "Gcell/s"
    from oracle import refshim
    # the unmodified reference when a copy is present (build container: /root/reference; GPU box: oracle/_ref), else the port
    assert line["cpu_baseline"]["kind"] == ("reference" if refshim.available() else "port")
    assert line["value"] > 0 and line["cpu_baseline"]["cores"] >= 1 and line["host"]["cpu_model"]
    if refshim.available():
        runs = line["extra"]["runs"]["128"]
        assert runs["solver_threading"]["uz_norm"] == runs["solver_default"]["uz_norm"] > 0      # the two reference solvers agree bit for bit
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] and "workload" in line["config"]


def test_merge_slabs_reassembles_single_gpu_schema(tmp_path):
    """h5lite.merge_slabs: per-slab files (attrs x0 / nxl, the plugin's multi-GPU output) -> one file."""
    from phonomena_b200.h5lite import H5Reader, H5Writer, merge_slabs
    rng = np.random.default_rng(5)
    nx, ny, N = 11, 4, 70
    full = {"ux": rng.standard_normal((nx - 1, ny, 1, N)), "uy": rng.standard_normal((nx, ny - 1, 1, N)),
            "uz": rng.standard_normal((nx, ny, 1, N))}
    dens = rng.uniform(1, 2, (nx, ny, 3))
    parts = []
    for r, (x0, nxl) in enumerate(((0, 4), (4, 4), (8, 3))):
        p = str(tmp_path / ("o.h5.rank%d" % r))
        with H5Writer(p) as w:
            w.attrs.update({"x": np.arange(nx, dtype=float), "dt": 1e-5, "steps": N, "x0": x0, "nxl": nxl,
                            "frames_written": N if r else N - 2, "prim_material": "GaAs"})
            w.create_dataset("density", dens[x0:x0 + nxl])
            for name, a in full.items():
                rows = min(x0 + nxl, a.shape[0]) - x0
                d = w.create_chunked(name, (rows,) + a.shape[1:])
                for t in range(N if r else N - 2):
                    w.write_frame(d, t, a[x0:x0 + rows, ..., t])
        parts.append(p)
    out = merge_slabs(parts[::-1], str(tmp_path / "o.h5"))
    r = H5Reader(out)
    assert r.attrs["x0"] == 0 and r.attrs["nxl"] == nx and r.attrs["frames_written"] == N - 2 and r.attrs["prim_material"] == "GaAs"
    assert np.array_equal(r.read("density"), dens)
    for name, a in full.items():
        assert r.shape(name) == a.shape
        for t in (0, 33, N - 3):
            assert np.array_equal(r.read(name, frame=t), a[..., t])
    with pytest.raises(ValueError, match="tile"):
        merge_slabs(parts[1:], str(tmp_path / "bad.h5"))


def test_h5lite_positional_frames_from_threads(tmp_path):
    """The plugin's writer path: space preallocated, frames reserved in any order and filled by concurrent
    positional writes of page-sized pieces; unused preallocated space is dropped at close."""
    import os
    from concurrent.futures import ThreadPoolExecutor
    from phonomena_b200.h5lite import H5Reader, H5Writer
    rng = np.random.default_rng(3)
    p = str(tmp_path / "par.h5")
    shape, frames = (40, 33, 1), 12
    data = {k: rng.standard_normal(shape + (frames,)) for k in ("ux", "uz")}
    w = H5Writer(p)
    w.attrs["dt"] = 0.5
    w.create_dataset("density", rng.standard_normal((5, 4, 3)))
    ds = {k: w.create_chunked(k, shape + (frames + 3,)) for k in data}      # three frames never written
    w.preallocate((frames + 3) * sum(d.frame_bytes + 8 for d in ds.values()) + (1 << 20))
    w.settle()
    with ThreadPoolExecutor(4) as pool:
        jobs = []
        for t in rng.permutation(frames):
            for k, d in ds.items():
                mv = memoryview(np.ascontiguousarray(data[k][..., t]).reshape(-1)).cast("B")
                pos = w.reserve_frame(d, int(t))
                for o in range(0, mv.nbytes, 4096):
                    jobs.append(pool.submit(w.pwrite, mv[o:o + 4096], pos + o))
        for j in jobs:
            j.result()
    with pytest.raises(ValueError):
        w.reserve_frame(ds["ux"], 0)                    # a frame is written once
    with pytest.raises(ValueError):
        w.reserve_frame(ds["ux"], frames + 3)           # outside the dataset
    w.close()
    r = H5Reader(p)
    for k in data:
        for t in range(frames):
            assert np.array_equal(r.read(k, frame=t), data[k][..., t]), (k, t)
    assert r.attrs["dt"] == 0.5 and r.shape("density") == (5, 4, 3)
    # the file ends where the metadata ends: the preallocated tail is gone
    assert os.path.getsize(p) < 2 * frames * sum(d.frame_bytes for d in ds.values()) + (1 << 16)


class _FakeRingEngine:
    """Stands in for _lib.Engine on the CPU: writer_start / writer_finish run the library's OWN ring + writer threads
    (phb_writer_selftest: a host producer pushes synthetic frames, element q of frame f = f * 1e6 + q, through the ring
    the device would fill) against the extents the plugin's Writer reserved."""

    def __init__(self, nx, ny, nz, produce, fail_fd=False):
        self.nx, self.ny, self.nz, self.x0, self.nxl = nx, ny, nz, 0, nx
        self.produce, self.fail_fd, self.args, self.aborted = produce, fail_fd, None, None

    def planes(self, comp):
        return self.nx - 1 if comp == 0 else self.nx

    def writer_start(self, fd, base, nbytes, stride, frames, nthreads=4, mmap=False, populate=False):
        self.args = (fd, list(base), list(nbytes), stride, frames, nthreads)
        self.mmap = mmap

    def writer_finish(self, timeout_ms=300000):
        from phonomena_b200 import _lib
        fd, base, nbytes, stride, frames, nthreads = self.args
        assert self.produce <= frames
        if self.fail_fd:
            fd = os.open(os.devnull, os.O_RDONLY)       # pwrite on a read-only descriptor: EBADF
        try:
            n = _lib.writer_selftest(fd, base, nbytes, stride, self.produce, slots=3, nthreads=nthreads, timeout_ms=2000,
                                     mmap=self.mmap and not self.fail_fd)
        except _lib.PhbError:
            self.writer_stats = (0, 0.0, 0.0)
            raise
        finally:
            if self.fail_fd:
                os.close(fd)
        self.writer_stats = (n, 0.0, 0.0)
        return self.writer_stats

    def record_abort(self, why=""):
        self.aborted = why


@pytest.mark.parametrize("full,fields,produce", [(False, ("ux", "uy", "uz"), 7), (False, ("uz",), 7),
                                                  (True, ("ux", "uy", "uz"), 7), (True, ("uy", "uz"), 4)])
def test_plugin_writer_native_threads_on_synthetic_frames(tmp_path, full, fields, produce):
    """The plugin's Writer (file layout, up-front reservation of every frame's extent, preallocation, record_fields
    subsets) + the library's native writer threads draining a ring: the file holds every produced frame in order, with
    the reference's dataset shapes (z extent 1 in surface mode); a run that stops early (cancel) leaves the missing
    frames unwritten, not garbage."""
    from phonomena_b200.h5lite import H5Reader
    from phonomena_b200.solver_b200 import Writer
    rng = np.random.default_rng(11)
    nx, ny, nz, frames = 9, 8, 5, 7
    eng = _FakeRingEngine(nx, ny, nz, produce)
    meta = {"attrs": {"dt": 0.25, "x": np.arange(nx, dtype=float)}, "density": rng.standard_normal((nx, ny, nz)), "elasticity": None}
    w = Writer(str(tmp_path / "w.h5"), eng, meta, frames, "full" if full else "surface", 1, ring=True, fields=fields)
    w.start()
    w.finish()
    w.finish()                                  # idempotent
    assert w.error is None and w.written == produce
    shp = {"ux": (nx - 1, ny), "uy": (nx, ny - 1), "uz": (nx, ny)}
    zext = {"ux": nz, "uy": nz, "uz": nz - 1}
    r = H5Reader(w.path)
    assert sorted(k for k in r.datasets if k.startswith("u")) == sorted(fields)
    assert r.attrs["frames_written"] == produce and r.attrs["record"] == ("full" if full else "surface")
    q0 = 0
    for k in fields:                            # ring order = ux, uy, uz as selected
        fshape = shp[k] + ((zext[k],) if full else (1,))
        n = int(np.prod(fshape))
        assert r.shape(k) == fshape + (frames,), (k, r.shape(k))
        for t in range(produce):
            exp = (t * 1e6 + q0 + np.arange(n, dtype=np.float64)).reshape(fshape)
            assert np.array_equal(r.read(k, frame=t), exp), (k, t)
        whole = r.read(k)
        assert not whole[..., produce:].any()   # frames never written read as zeros
        q0 += n
    assert np.array_equal(r.read("density"), meta["density"])


def test_plugin_writer_error_surfaces_and_abort_is_quiet(tmp_path):
    """A write error inside the native threads (here EBADF) comes back from finish() as an exception carrying the errno
    text, and the file is still closed as a valid (empty) recording; abort() -- what init()-again and __del__ use --
    never raises."""
    from phonomena_b200 import _lib
    from phonomena_b200.h5lite import H5Reader
    from phonomena_b200.solver_b200 import Writer
    meta = {"attrs": {"dt": 0.25}, "density": np.zeros((6, 5, 4)), "elasticity": None}
    eng = _FakeRingEngine(6, 5, 4, 5, fail_fd=True)
    w = Writer(str(tmp_path / "e.h5"), eng, meta, 5, "surface", 1, fields=("uz",))
    w.start()
    with pytest.raises(_lib.PhbError, match="pwrite"):
        w.finish()
    r = H5Reader(w.path)
    assert r.attrs["frames_written"] == 0 and r.shape("uz") == (6, 5, 1, 5)
    eng2 = _FakeRingEngine(6, 5, 4, 5, fail_fd=True)
    w2 = Writer(str(tmp_path / "a.h5"), eng2, meta, 5, "surface", 1, fields=("uz",))
    w2.start()
    w2.abort("re-init")
    assert eng2.aborted == "re-init" and w2.finished and H5Reader(w2.path).attrs["frames_written"] == 0


def test_native_ring_flow_control_and_threads(tmp_path):
    """phb_writer_selftest directly: more frames than ring slots (the producer has to wait for in-order releases), 1..4
    writer threads, interleaved component extents."""
    from phonomena_b200 import _lib
    for nthreads, mm in ((1, False), (2, False), (4, False), (3, True)):
        p = str(tmp_path / ("ring%d.bin" % nthreads))
        fd = os.open(p, os.O_RDWR | os.O_CREAT | os.O_TRUNC, 0o644)
        nb = [24 * 8, 40 * 8]
        stride, frames = sum(nb) + 64, 50
        base = [5000, 5000 + nb[0]]          # not page-aligned: the mapped path maps from the page below
        if mm:
            os.posix_fallocate(fd, 0, base[0] + stride * frames)
        assert _lib.writer_selftest(fd, base, nb, stride, frames, slots=3, nthreads=nthreads, mmap=mm) == frames
        os.close(fd)
        raw = np.fromfile(p, dtype="<f8")
        for f in (0, 1, 17, 49):
            a = raw[(base[0] + f * stride) // 8:][:24]
            b = raw[(base[1] + f * stride) // 8:][:40]
            assert np.array_equal(a, f * 1e6 + np.arange(24)) and np.array_equal(b, f * 1e6 + 24 + np.arange(40)), (nthreads, f)


def test_h5writer_dataset_from_pieces(tmp_path):
    """H5Writer.create_dataset_from: a contiguous dataset streamed in plane blocks equals the one written at once, and a
    short iterator is an error, not a silently truncated dataset."""
    from phonomena_b200.h5lite import H5Reader, H5Writer
    from tests import h5check
    rng = np.random.default_rng(4)
    a = rng.standard_normal((37, 5, 6))
    p = str(tmp_path / "pieces.h5")
    with H5Writer(p) as w:
        w.create_dataset_from("density", a.shape, (a[q:q + 16] for q in range(0, 37, 16)))
        w.create_dataset("other", a[:3])
        w.attrs["dt"] = 1.0
    r = H5Reader(p)
    assert np.array_equal(r.read("density"), a) and np.array_equal(r.read("other"), a[:3])
    h5check.validate(p)
    w = H5Writer(str(tmp_path / "short.h5"))
    with pytest.raises(ValueError, match="pieces"):
        w.create_dataset_from("density", a.shape, (a[q:q + 16] for q in range(0, 32, 16)))
