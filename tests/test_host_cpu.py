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
    out, idx = analysis.trim_trailing_zeros(d1.copy())
    assert out.size == len(idx) <= d1.size


def test_bench_reference_arm_contract_line():
    """`bench.py --impl reference` (CPU only): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["metric"] == "Gcell-updates/s" and line["unit"] == "Gcell/s"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
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
    """Stands in for _lib.Engine on the CPU: hands the plugin's Writer a scripted sequence of recorder frames
    (the real one returns views into the pinned ring; record_next / record_release keep the same contract)."""

    def __init__(self, nx, ny, nz, frames, fields, full, rng):
        self.nx, self.ny, self.nz, self.x0, self.nxl = nx, ny, nz, 0, nx
        shp = {"ux": (nx - 1, ny), "uy": (nx, ny - 1), "uz": (nx, ny)}
        zext = {"ux": nz, "uy": nz, "uz": nz - 1}
        self.data = [{k: rng.standard_normal(shp[k] + ((zext[k],) if full else ())) for k in fields} for _ in range(frames)]
        self.next, self.held, self.released = 0, False, 0

    def planes(self, comp):
        return self.nx - 1 if comp == 0 else self.nx

    def record_next(self, timeout_ms=0):
        assert not self.held, "record_next called again before record_release"
        if self.next >= len(self.data):
            return None
        self.held = True
        self.next += 1
        return self.next - 1, self.data[self.next - 1]

    def record_release(self):
        assert self.held
        self.held = False
        self.released += 1


@pytest.mark.parametrize("full,fields,parallel", [(False, ("ux", "uy", "uz"), True), (False, ("uz",), False),
                                                   (True, ("ux", "uy", "uz"), True), (True, ("uy", "uz"), False)])
def test_plugin_writer_thread_on_scripted_frames(tmp_path, full, fields, parallel):
    """The plugin's Writer (ring consumer thread, preallocation, piecewise parallel positional writes above
    PARALLEL_MIN_BYTES, one write below it, record_fields subsets) against a scripted frame source: the file holds
    every frame in order, with the reference's dataset shapes (z extent 1 in surface mode)."""
    from phonomena_b200.h5lite import H5Reader
    from phonomena_b200.solver_b200 import Writer
    rng = np.random.default_rng(11)
    nx, ny, nz, frames = 9, 8, 5, 7
    eng = _FakeRingEngine(nx, ny, nz, frames, fields, full, rng)
    meta = {"attrs": {"dt": 0.25, "x": np.arange(nx, dtype=float)}, "density": rng.standard_normal((nx, ny, nz)), "elasticity": None}
    w = Writer(str(tmp_path / "w.h5"), eng, meta, frames, "full" if full else "surface", 1, ring=True, fields=fields)
    if parallel:
        w.PARALLEL_MIN_BYTES = 64          # force the multi-piece path on these tiny frames
    w.start()
    w.finish()
    assert w.error is None and w.written == frames and eng.released == frames
    r = H5Reader(w.path)
    assert sorted(k for k in r.datasets if k.startswith("u")) == sorted(fields)
    assert r.attrs["frames_written"] == frames and r.attrs["record"] == ("full" if full else "surface")
    for k in fields:
        exp_shape = eng.data[0][k].shape + (() if full else (1,)) + (frames,)
        assert r.shape(k) == exp_shape, (k, r.shape(k), exp_shape)
        for t in range(frames):
            got = r.read(k, frame=t)
            assert np.array_equal(got if full else got[..., 0], eng.data[t][k]), (k, t)
    assert np.array_equal(r.read("density"), meta["density"])
