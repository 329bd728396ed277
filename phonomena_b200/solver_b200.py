"""B200 solver plugin for Phonomena -- drop-in for the solver interface of
``phonomena/simulation/solvers/solver_*.py`` (reference: ``base_solver.BaseSolver``,
base_solver.py:172-292, and ``solver_default.Solver``).

Same public surface as the reference plugins:

    s = Solver()                      # zero-arg constructor (common.py:88-91)
    s.name, s.description, s.cfg, s.file, s.running, s.logger
    s.init(grid, material, steps)     # base_solver.py:194-222
    s.run(signals=...)                # base_solver.py:224-280  (status / progress signals)
    s.cancel()                        # base_solver.py:282-284
    s.test()                          # base_solver.py:286-292

What changes is what happens inside: the time loop is not NumPy on the host but the sm_100a
kernels of libphb200.so (include/phb200.h) driven through ctypes; the per-cell material is
generated on the device from the inclusion list; the surface plane is recorded by the device
into a pinned host ring and flushed to HDF5 by a writer thread.  No CPU fallback: if the
library or the GPU is missing, ``init`` raises.

To install into a Phonomena checkout, drop a three-line ``solver_b200.py`` into
``phonomena/simulation/solvers/``:

    from phonomena_b200.solver_b200 import Solver   # noqa: F401
"""
from __future__ import annotations

import atexit
import copy
import json
import logging
import os
import tempfile
import threading
import time

import numpy as np

from . import _lib, hostmath as hm
from .h5lite import H5Writer, merge_slabs

logger = logging.getLogger(__name__)
_live_solvers = None      # weak set of Solver instances, closed at interpreter exit before the engines are (see _lib._close_all)

# module-level cfg, merged with the base defaults like the reference plugins do
# (solver_numba.py:8-13,22).  All values are JSON-serialisable (common.saveSettings dumps them).
cfg = {
    "precision": "fp64",        # "fp64" | "fp32": storage and arithmetic type on the device
    "arith": "fast",            # "fast" (<=1e-12 of the reference in fp64) | "exact" (bit-identical, slower) |
                                # "compensated" (state u, u-u_old: 4-8x less fp32 round-off drift on long runs)
    "device": 0,                # CUDA device ordinal (one process per GPU; slabs via torchrun, see slab_from_env)
    "record": "auto",           # "full": whole fields per step, the reference's schema (base_solver.py:105-133) | "surface": ux, uy,
                                # uz at z-index 0 only (datasets with z extent 1) | "off" | "auto": full while a frame is at most
                                # AUTO_FULL_MAX_FRAME_BYTES (the grids the reference's GUI handles), else surface
    "record_every": 1,
    "record_fields": ["ux", "uy", "uz"],   # which displacement components are recorded (the reference writes all three)
    "chunk_steps": 50,          # steps enqueued per library call (cancel / progress granularity)
    "kernel": "auto",
    "material": "inclusions",   # "inclusions": primary/secondary + inclusion list, filled on the device (what Material.update
                                # builds) | "arrays": take material.C / material.P as they are (any <= 15 distinct cells)
    "record_stride": [1, 1, 1], # record = "full": keep every sx-th plane / sy-th row / sz-th level (decimated volume snapshots of big grids)
    "merge_slabs": True,        # multi-GPU: concatenate the per-slab files into `file` after the run (rank 0)
    "bc_y": "absorbing",        # "absorbing": Mur faces at y = 0 / y = -1 (the reference) | "periodic": the reference's archived
                                # apply_T_pbc / apply_u_pbc stubs (zero Bloch phase) in their place | "bloch": the
                                # same with a phase, u(y + L) = u(y) exp(i bloch_phase): complex field, two device contexts
    "bloch_phase": 0.0,         # radians (bc_y = "bloch")
    "probes": [],               # [{"u": "uz", "y": j, "z": k}, ...]: (x, t) lines kept on the device for Solver.spectrum()
}


class _DummySignal:
    def emit(self, *a, **k):
        pass


class _DummySignals:
    """Stand-in for gui.worker.WorkerSignals() when run() gets no `signals` (base_solver.py:228-230)."""
    def __init__(self):
        self.status = self.progress = self.error = self.finished = _DummySignal()


class Writer:
    """Replaces base_solver.Writer (:72-169): same file schema (App. C), fed from the device ring.

    surface mode: datasets ux (Nx-1,Ny,1,frames), uy (Nx,Ny-1,1,frames), uz (Nx,Ny,1,frames),
    chunk = one frame, so `u[:, :, 0, t]` of the consumers (h5py2gif.py:24,44; analysis.py:59-66)
    works unchanged.  full mode: the reference's full 4-D datasets.

    This class owns the file and its layout (h5lite); the frames themselves never pass through Python:
    the chunk extents of all frames are reserved up front and native threads inside libphb200.so
    (phb_writer_start, csrc/rec_ring.h) pwrite() every recorded frame from the pinned ring straight to
    its extent.  A write error (ENOSPC ...) aborts the recording, which makes the stepping call fail with
    the errno text -- the run loop cannot hang on a dead writer (reference: queue.get(timeout=120),
    join(300), base_solver.py:89-92,148,274)."""

    WRITE_THREADS = 4      # native threads, each writes whole frames
    PREALLOCATE_MAX_BYTES = 8 << 30

    def __init__(self, path, engine, meta, frames, mode, record_every, ring=True, fields=("ux", "uy", "uz"), stride=(1, 1, 1)):
        self.path, self.e, self.mode, self.frames = path, engine, mode, frames
        self.stride = tuple(stride) if mode == "full" else (1, 1, 1)
        self.ring = ring or mode == "surface"      # frames arrive through the device / pinned ring (else: put_full)
        self.h5 = H5Writer(path)
        self.h5.attrs.update(meta["attrs"])
        self.h5.attrs["record_every"] = int(record_every)
        self.h5.attrs["record"] = mode
        if callable(meta["density"]):          # (shape, piece iterator): written plane block by plane block
            dshape, pieces = meta["density"]()
            self.h5.create_dataset_from("density", dshape, pieces)
        else:
            self.h5.create_dataset("density", meta["density"])
        if meta.get("elasticity") is not None:
            self.h5.create_dataset("elasticity", meta["elasticity"])
        ny, nz = engine.ny, engine.nz
        zext = 1 if mode == "surface" else None
        # x extents are the planes this context owns (the whole grid, or one slab: attrs x0 / nxl)
        cd = lambda n, s: -(-n // s)
        sx, sy, sz = self.stride
        shapes = {"ux": (cd(engine.planes(0), sx), cd(ny, sy), zext or cd(nz, sz), frames),
                  "uy": (cd(engine.planes(1), sx), cd(ny - 1, sy), zext or cd(nz, sz), frames),
                  "uz": (cd(engine.planes(2), sx), cd(ny, sy), zext or cd(nz - 1, sz), frames)}
        self.ds = {k: self.h5.create_chunked(k, shapes[k]) for k in ("ux", "uy", "uz") if k in fields}
        # The static datasets (density alone is Nx*Ny*Nz doubles) are pushed out now, in init(), like the reference's
        # Writer.init (base_solver.py:105-133): otherwise the kernel's dirty-page writeback of them throttles the
        # frame appends of the stepping loop.
        self._layout, self._mapped = None, False
        if self.ring and frames > 0:
            start = self.h5._end
            base, stride = self.h5.reserve_frames(list(self.ds.values()), frames)
            self._layout = (base, [d.frame_bytes for d in self.ds.values()], stride)
            # allocated extents can be mapped: the native threads then copy into the page cache without sharing the
            # file's inode lock (pwrite serialises on it); unallocated ones are written with pwrite (ENOSPC -> error, not SIGBUS)
            self._mapped = stride * frames <= self.PREALLOCATE_MAX_BYTES and bool(self.h5.preallocate(self.h5._end - start, start))
        self.h5.settle()
        self.written = 0
        self.wait_seconds = self.write_seconds = 0.0     # writer threads: waiting for frames / writing them
        self.error = None
        self.started = self.finished = False

    def start(self):
        if self._layout is None:
            return
        base, nbytes, stride = self._layout
        # populate: the allocated pages are entered into the mapping now (no zeroing, ~0.2 us per page), so the stepping
        # loop's copies take no page faults (first-touch faults on one file serialise on its page-cache lock: 6.8 GB/s
        # for any number of threads, measured)
        self.e.writer_start(self.h5.fd, base, nbytes, stride, self.frames, self.WRITE_THREADS, mmap=self._mapped, populate=self._mapped)
        self.started = True

    # full mode without a ring: called synchronously by the run loop
    def put_full(self, fields):
        sx, sy, sz = self.stride
        for name, a in zip(("ux", "uy", "uz"), fields):
            if name in self.ds:
                self.h5.write_frame(self.ds[name], self.written, a[::sx, ::sy, ::sz])
        self.written += 1

    def finish(self, timeout=300.0):
        """Drain, join the native threads, close the file.  Raises what the writer threads hit."""
        if self.finished:
            return
        self.finished = True
        if self.started:
            try:
                self.e.writer_finish(int(timeout * 1000))
            except Exception as exc:
                self.error = exc
            self.written, self.wait_seconds, self.write_seconds = getattr(self.e, "writer_stats", (0, 0.0, 0.0))
            self.h5.keep_frames(self.written)
        self.h5.attrs["frames_written"] = int(self.written)
        try:
            self.h5.close()
        except OSError as exc:          # the error that stopped the writer threads (ENOSPC, ...) usually hits the metadata too
            if self.error is None:
                self.error = exc
        if self.error is not None:
            raise self.error

    def abort(self, why="writer abandoned"):
        """Stop without raising (init() called again, or the solver is being destroyed): the file is closed with the
        frames written so far, the native threads are joined before the engine goes away."""
        if self.finished:
            return
        if self.started:
            self.e.record_abort(why)
        try:
            self.finish(timeout=5.0)
        except Exception:
            pass


class Solver:
    AUTO_FULL_MAX_FRAME_BYTES = 64 << 20     # record = "auto": whole fields up to this frame size (about 140^3 points), surface planes above
    FULL_RING_MAX_FRAME_BYTES = 1 << 30      # record = "full": frames up to this size stream through the recorder ring
    FULL_RING_BYTES = 2 << 30                # pinned + device staging ring budget for them (>= 2 slots)

    def __init__(self):
        self.name = "b200"
        self.description = ("<p>FDTD time stepping on an NVIDIA B200 (sm_100a): fused stress + displacement "
                            "kernel fed by TMA, absorbing / free-surface boundaries, on-device surface recording. "
                            "cfg: precision fp64|fp32, arith fast|exact, record surface|full|off.</p>")
        with tempfile.NamedTemporaryFile() as f:
            self.file = os.path.realpath(f.name)
        self.running = threading.Event()
        self.logger = logger
        # combine module cfg and the base-class defaults (base_solver.py:190-192)
        self.cfg = {**cfg, "wave": "ricker", "wave_args": {"f": 100}, "write_mode": "process"}
        self.engine = None
        self.writer = None
        self.stats = {}
        global _live_solvers
        if _live_solvers is None:
            import weakref
            _live_solvers = weakref.WeakSet()
            atexit.register(_close_solvers)
        _live_solvers.add(self)

    # ------------------------------------------------------------------------------------
    def init(self, grid, material, steps):
        """Prepare a run: copies what it needs (the caller's objects are not mutated and may be
        freed), rebuilds the mesh like the reference does, uploads everything, zeroes the fields."""
        self._close_engine()       # also stops a writer left over from a previous init() (base_solver.py:209-211)
        self._ran = False
        t_init = time.time()
        c = self.cfg
        self.t = int(copy.deepcopy(steps))
        logger.info("Initializing %s with settings: %s", self.name, c)

        # -- mesh: the reference deep-copies the grid and re-runs buildMesh (base_solver.py:199-206).
        #    Field arrays are not needed here, so only the mesh description is copied.
        g = _mesh_copy(grid)
        if hasattr(g, "buildMesh") and getattr(g, "targets", None) is not None and hasattr(g, "spacing_fn"):
            g.buildMesh()
        x, y, z = (np.array(a, np.float64) for a in (g.x, g.y, g.z))
        si = getattr(g, "SI_conversion", 1)
        fdx, fdy, fdz, sdx, sdy, sdz = hm.spacings(x, y, z, si)

        # -- material: m.update() of the reference works on the material's OWN grid copy
        #    (material.py:49-50, App. B #11) -> inclusion list and mesh lines come from material.grid
        mg = getattr(material, "grid", None) or g
        mx, my, mz = (np.array(a, np.float64) for a in (mg.x, mg.y, mg.z))
        if (mx.size, my.size, mz.size) != (x.size, y.size, z.size):
            raise ValueError("material.grid has %s points, solver grid has %s" % ((mx.size, my.size, mz.size), (x.size, y.size, z.size)))
        prim = {"c": np.array(material.primary["c"], np.float64), "p": float(material.primary["p"]),
                "name": material.primary.get("name", "primary")}
        sec = {"c": np.array(material.secondary["c"], np.float64), "p": float(material.secondary["p"]),
               "name": material.secondary.get("name", "secondary")}
        msi = getattr(mg, "SI_conversion", 1)
        mfd = hm.spacings(mx, my, mz, msi)
        dt = hm.cfl_dt(mfd[0], mfd[1], mfd[2], material.c_max, prim, sec, msi)
        targets = _targets_array(getattr(mg, "targets", None))

        rec_mode = c["record"] if c.get("write_mode", "off") != "off" else "off"
        x0, nxl, rank, nranks = slab_from_env(x.size) if c.get("slabs_from_env") else (0, x.size, 0, 1)
        if rec_mode == "auto":
            # the reference writes the whole fields every step; that is what its consumers index (z_index > 0 in the GUI's
            # spectrum tab).  Keep that schema wherever it is affordable, fall back to the surface planes on big grids.
            nf = len([k for k in ("ux", "uy", "uz") if k in c.get("record_fields", ("ux", "uy", "uz"))]) or 1
            rec_mode = "full" if 8 * x.size * y.size * z.size * nf <= self.AUTO_FULL_MAX_FRAME_BYTES else "surface"
        if rec_mode not in ("surface", "full", "off"):
            raise ValueError("cfg['record'] must be 'auto', 'surface', 'full' or 'off' (got %r)" % (c["record"],))
        rec_mask, ring_slots = 0, 32
        fields = tuple(k for k in ("ux", "uy", "uz") if k in c.get("record_fields", ("ux", "uy", "uz")))
        if rec_mode != "off" and not fields:
            raise ValueError("record_fields selects none of ux, uy, uz")
        fmask = sum(b for k, b in (("ux", _lib.REC_UX), ("uy", _lib.REC_UY), ("uz", _lib.REC_UZ)) if k in fields)
        if rec_mode == "surface":
            rec_mask = fmask
        elif rec_mode == "full":
            # whole arrays every recorded step (the reference's Writer, base_solver.py:97-100,135-160): when a frame is
            # small enough they go through the same device ring -> pinned ring -> writer thread as the surface planes,
            # so the stepping loop never waits for a read-back; big grids keep the synchronous get_fields path
            # (cfg["record_stride"] decimates the volume: 512^3 at stride 4 is a 50 MB frame instead of 3.2 GB)
            stride = tuple(max(1, int(v)) for v in (list(c.get("record_stride") or [1, 1, 1]) + [1, 1, 1])[:3])
            if x0 % stride[0]:
                raise ValueError("record_stride[0] = %d does not divide this slab's origin x0 = %d" % (stride[0], x0))
            fbytes = 8 * -(-nxl // stride[0]) * -(-y.size // stride[1]) * -(-z.size // stride[2]) * len(fields)
            if fbytes <= self.FULL_RING_MAX_FRAME_BYTES:
                rec_mask = fmask | _lib.REC_FULL
                ring_slots = int(max(2, min(32, self.FULL_RING_BYTES // max(1, fbytes))))
        else:
            stride = (1, 1, 1)
        if rec_mode != "full":
            stride = (1, 1, 1)
        self._stride = stride
        self._full_ring = bool(rec_mask & _lib.REC_FULL)
        e = _lib.Engine(x.size, y.size, z.size, dt, d2=dt ** 2,
                        dtype={"fp64": "f64", "fp32": "f32"}[c["precision"]], arith=c["arith"],
                        device=int(c["device"]), x0=x0, nxl=nxl, kernel=c.get("kernel", "auto"),
                        record_mask=rec_mask, record_every=int(c["record_every"]), ring_slots=ring_slots,
                        bc_y={"bloch": "periodic"}.get(c.get("bc_y", "absorbing"), c.get("bc_y", "absorbing")), record_stride=stride)
        self.engine = e
        self.engine_imag = None
        if c.get("bc_y") == "bloch":
            if nranks > 1:
                raise ValueError("bc_y = 'bloch' runs on one GPU")
            # the imaginary part: same grid / material / faces, no source, no recorder; paired below
            self.engine_imag = _lib.Engine(x.size, y.size, z.size, dt, d2=dt ** 2, dtype={"fp64": "f64", "fp32": "f32"}[c["precision"]],
                                           arith=c["arith"], device=int(c["device"]), kernel=c.get("kernel", "auto"), bc_y="periodic")
        if nranks > 1:
            # one process per GPU: fused NVLink halo push (CUDA IPC handles all-gathered over any host channel,
            # default torch.distributed), NCCL send/recv as fallback
            self.halo = e.connect(rank, nranks, self.allgather, self.broadcast, mode=c.get("halo"))
        e.set_spacing(fdx, fdy, fdz, sdx, sdy, sdz)
        dense = c.get("material", "inclusions") == "arrays"
        if dense:
            # the reference's own per-cell arrays (material.py:48-63): whatever the caller put there
            Cd, Pd = np.asarray(material.C), np.asarray(material.P)
            if Pd.shape != (x.size, y.size, z.size) or Cd.shape != Pd.shape + (6, 6):
                raise ValueError("material.C / material.P have shapes %s / %s for a %s grid" % (Cd.shape, Pd.shape, (x.size, y.size, z.size)))
            sl = slice(x0, x0 + e.id_planes())
            e.set_material_dense(Cd[sl], Pd[sl])
            cm = {"c": np.array(Cd[0, 0, 0], np.float64), "p": float(Pd[0, 0, 0])}
            P_out = np.array(Pd[x0:x0 + nxl], np.float64) if rec_mode != "off" else None
            C_out = Cd[x0:x0 + nxl]
        else:
            e.set_material_table([prim["c"], sec["c"]], [prim["p"], sec["p"]])
            e.gen_material_ids(targets, mx, my, mz)
            ids = e.get_material_ids() if (rec_mode != "off" or x0 == 0) else None
            corner = int(ids[0, 0, 0]) if (ids is not None and x0 == 0) else 0
            if nranks > 1:      # the Mur coefficients come from the corner cell (0,0,0), which rank 0 owns
                corner = int(self.broadcast(corner if rank == 0 else None))
            cm = sec if corner else prim
        e.set_abc(hm.abc_coefficients(cm["c"], cm["p"], dt, fdx, fdy, fdz, sdx, sdy, sdz))
        if self.engine_imag is not None:
            ei = self.engine_imag
            ei.set_spacing(fdx, fdy, fdz, sdx, sdy, sdz)
            if dense:
                ei.set_material_dense(Cd[sl], Pd[sl])
            else:
                ei.set_material_table([prim["c"], sec["c"]], [prim["p"], sec["p"]])
                ei.gen_material_ids(targets, mx, my, mz)
            ei.set_abc(hm.abc_coefficients(cm["c"], cm["p"], dt, fdx, fdy, fdz, sdx, sdy, sdz))
            e.bloch_pair(ei, float(c.get("bloch_phase", 0.0)))
        self.dt, self._x0 = dt, x0
        self._x, self._fdx, self._nranks = x, fdx, nranks
        # line probes: the (x, t) matrices analysis.spectrum would re-read from the file stay in HBM
        self._probes = {}
        for pr in c.get("probes") or []:
            key = (pr["u"], int(pr["y"]), int(pr.get("z", 0)))
            self._probes[key] = e.probe_add(key[0], key[1], key[2], max(1, self.t // int(c["record_every"])))
        self._wave, self._wave_args = c["wave"], dict(c["wave_args"])

        self.writer = None
        if rec_mode != "off":
            frames = self.t // int(c["record_every"])
            if dense:
                P = P_out
            else:
                # density = table[id]: produced plane block by plane block while it is written (a 512^3 grid would otherwise
                # need a 1 GB float64 array plus a bytes copy of it: most of init()'s time)
                lut = np.array([prim["p"], sec["p"]], np.float64)
                ids_d = ids[::stride[0], ::stride[1], ::stride[2]] if stride != (1, 1, 1) else ids

                def P(ids_d=ids_d, lut=lut):
                    return ids_d.shape, (lut[ids_d[q:q + 16]] for q in range(0, ids_d.shape[0], 16))
            attrs = {"x": x, "y": y, "z": z,
                     "sdx": sdx.reshape(-1, 1, 1), "sdy": sdy.reshape(1, -1, 1), "sdz": sdz.reshape(1, 1, -1),
                     "fdx": fdx.reshape(-1, 1, 1), "fdy": fdy.reshape(1, -1, 1), "fdz": fdz.reshape(1, 1, -1),
                     "steps": int(self.t), "dt": float(dt), "prim_material": prim["name"], "sec_material": sec["name"],
                     "solver_cfg": json.dumps(c), "x0": int(x0), "nxl": int(nxl)}
            if stride != (1, 1, 1):
                # decimated volume: the file describes the decimated mesh (consumers index x / fdx by u.shape), the full
                # mesh lines are kept beside it
                sx_, sy_, sz_ = stride
                xs, ys, zs = x[::sx_], y[::sy_], z[::sz_]
                attrs.update({"x_full": x, "y_full": y, "z_full": z, "record_stride": np.array(stride, np.float64), "x": xs, "y": ys, "z": zs})
                for key, line, shp in (("fdx", xs, (-1, 1, 1)), ("fdy", ys, (1, -1, 1)), ("fdz", zs, (1, 1, -1))):
                    fd = np.diff(line) * si
                    attrs[key] = fd.reshape(shp)
                    attrs["s" + key[1:]] = (0.5 * (fd[1:] + fd[:-1])).reshape(shp)
                if dense:
                    P = np.ascontiguousarray(P[::sx_, ::sy_, ::sz_])
            meta = {"attrs": attrs, "density": P, "elasticity": None}
            n_density = (ids[::stride[0], ::stride[1], ::stride[2]].size if not dense else P.size)
            if rec_mode == "full" and n_density <= 1 << 22:      # the reference's `elasticity` is 288 B/cell
                Cfull = np.array(C_out, np.float64) if dense else np.where((ids == 1)[..., None, None], sec["c"], prim["c"])
                meta["elasticity"] = np.ascontiguousarray(Cfull[::stride[0], ::stride[1], ::stride[2]])
            path = self.file if nranks == 1 else "%s.rank%d" % (self.file, rank)      # one file per slab
            self.writer = Writer(path, e, meta, frames, rec_mode, int(c["record_every"]), ring=self._full_ring, fields=fields, stride=stride)
            self.writer.start()
        self._rec_mode = rec_mode
        self.init_seconds = time.time() - t_init

    # ------------------------------------------------------------------------------------
    def run(self, *args, **kwargs):
        signals = kwargs.get("signals") or _DummySignals()
        if self.engine is None:
            raise RuntimeError("init() must be called before run()")
        if getattr(self, "_ran", False):
            # the reference's second run() dies on its finished writer ("Process not started" / closed queue); here the
            # recorder would wait for a consumer that no longer exists -- refuse instead
            raise RuntimeError("run() was already called for this init(); call init() again")
        self._ran = True
        e, c = self.engine, self.cfg
        signals.status.emit("Solver starting..")
        self.running.set()
        signals.status.emit("Running simulation.")
        stime = time.time()
        signals.progress.emit(0)
        progress = 0
        chunk = max(1, int(c["chunk_steps"]))
        every = int(c["record_every"])
        done = 0
        l0 = e.launch_count
        try:
            while done < self.t:
                if not self.running.is_set():
                    self.logger.warning("Simulation cancelled.")
                    break
                n = min(chunk, self.t - done)
                if self._rec_mode == "full" and not self._full_ring:
                    n = min(n, every - (done % every))
                # the source samples of this chunk: evaluated on the host with the reference's
                # expression and copied to the device inside the loop (base_solver.py:251)
                e.set_source_table(hm.source_table(self._wave, n, self.dt, self._wave_args, start=done))
                try:
                    e.run(n)
                except _lib.PhbCancelled:       # cancel() reached the library between two steps of this chunk
                    done = e.steps_done
                    self.logger.warning("Simulation cancelled.")
                    break
                done += n
                if self._rec_mode == "full" and not self._full_ring:
                    if done % every == 0:
                        self.writer.put_full(e.get_fields())
                elif self._rec_mode not in ("surface", "full"):
                    e.sync()
                if progress < 99:
                    progress = min(99, int((done / self.t) * 100))
                    signals.progress.emit(progress)
            e.sync()
            self._loop_seconds = time.time() - stime
        finally:
            self.running.clear()
            if self.writer is not None:
                t_fin = time.time()
                self.writer.finish()
                self._finish_seconds = time.time() - t_fin
        if self._nranks > 1 and self.writer is not None and c.get("merge_slabs", True):
            # every slab file is closed (the all-gather doubles as the barrier); rank 0 concatenates along x
            parts = self.allgather(self.writer.path)
            if self._x0 == 0:
                merge_slabs(parts, self.file)
                for p in parts:
                    os.remove(p)
            self.allgather(None)
        etime = time.time() - stime
        cells = e.nx * e.ny * e.nz
        self.stats = {"steps": done, "seconds": etime, "gcells_per_s": cells * done / etime / 1e9 if etime > 0 else 0.0,
                      "launches": e.launch_count - l0, "kernel": e.info()["kernel"],
                      "loop_seconds": getattr(self, "_loop_seconds", 0.0), "init_seconds": getattr(self, "init_seconds", 0.0),
                      "writer_finish_seconds": getattr(self, "_finish_seconds", 0.0),
                      "writer_wait_seconds": self.writer.wait_seconds if self.writer is not None else 0.0,
                      "writer_write_seconds": self.writer.write_seconds if self.writer is not None else 0.0}
        signals.status.emit("Simulation finished in {:.2f}s ({:.2f} Gcell/s).".format(etime, self.stats["gcells_per_s"]))
        signals.progress.emit(100)

    def cancel(self):
        if self.running.is_set():
            self.running.clear()
            e = self.engine
            if e is not None:
                e.cancel()      # the chunk in flight stops after its current step (also when it waits for the recorder ring)

    def allgather(self, obj):
        """All-gather a small Python object over the ranks (multi-GPU runs); see broadcast()."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("multi-GPU run: initialise torch.distributed (torchrun) or override Solver.allgather")
        out = [None] * dist.get_world_size()
        dist.all_gather_object(out, obj)
        return out

    def broadcast(self, obj):
        """Broadcast a small Python object from rank 0 (multi-GPU runs).  Default channel:
        torch.distributed, which the launcher (torchrun) has initialised; override to use another."""
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("multi-GPU run: initialise torch.distributed (torchrun) or override Solver.broadcast")
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def test(self):
        """base_solver.py:286-292: 10 steps on the reference's TestDefaults grid (needs the reference
        package on sys.path, as it is when this file sits in its solvers directory)."""
        from simulation import base_solver     # reference module
        self.init(grid=base_solver.TestDefaults.g, material=base_solver.TestDefaults.m, steps=10)
        self.run()

    def spectrum(self, u_id, z_index, y_index, x_index=None):
        """simulation/analysis.py:44-96 without the file: (x, f, dft) of the Hann-windowed 1-D (t) or
        2-D (x, t) transform of u_id[:, y_index, z_index, :], computed on the device from the line
        probe registered through cfg["probes"].  Multi-GPU: every rank returns the full result."""
        key = (u_id, int(y_index), int(z_index))
        if self.engine is None or key not in self._probes:
            raise KeyError("no probe for %s: add {'u': %r, 'y': %d, 'z': %d} to cfg['probes'] before init()" % ((key,) + key))
        e, pid = self.engine, self._probes[key]
        rows, N = e.probe_shape(pid)
        nxt = self._x.size - (1 if u_id == "ux" else 0)
        Nf = N // 2
        x = hm.nonlinspace(self._fdx) if u_id == "ux" else np.array(self._x)
        f = np.fft.fftfreq(N, d=self.dt * int(self.cfg["record_every"]))[:Nf]
        window = np.hanning(N)
        if x_index is None:
            part = e.probe_dft_xt(pid, window, Nf, nxt)
            if self._nranks > 1:
                part = sum(self.allgather(part))
            return x, f, np.abs(part) / np.sqrt(float(nxt) * N)
        xi = int(x_index) % nxt
        mine = self._x0 <= xi < self._x0 + rows
        line = e.probe_dft_t(pid, window, Nf, xi - self._x0, 1)[0] if mine else None
        if self._nranks > 1:
            line = next(a for a in self.allgather(line) if a is not None)
        return x, f, np.abs(line) / np.sqrt(float(N))

    # -- helpers for tests / scripts -----------------------------------------------------------
    def fields(self):
        """(ux, uy, uz) of the current state in the reference's shapes (bc_y = "bloch": the real part)."""
        return self.engine.get_fields()

    def fields_imag(self):
        """bc_y = "bloch": the imaginary part of the current state."""
        if getattr(self, "engine_imag", None) is None:
            raise RuntimeError("no imaginary part: cfg['bc_y'] is not 'bloch'")
        self.engine.sync()
        return self.engine_imag.get_fields()

    def _close_engine(self):
        if self.writer is not None:
            self.writer.abort("solver re-initialised or closed")     # joins the native writer threads, closes the file
            self.writer = None
        if self.engine is not None:
            self.engine.close()
            self.engine = None
        if getattr(self, "engine_imag", None) is not None:
            self.engine_imag.close()
            self.engine_imag = None

    def __del__(self):
        try:
            self._close_engine()
        except Exception:
            pass


def _close_solvers():
    for s in list(_live_solvers or ()):
        try:
            s._close_engine()
        except Exception:
            pass


# ----------------------------------------------------------------------------------------------
def _mesh_copy(grid):
    """Deep copy of the mesh description without the 18 field arrays Grid.update allocates."""
    skip = {"ux", "uy", "uz", "T1", "T2", "T3", "T4", "T5", "T6"}
    if not hasattr(grid, "__dict__"):
        return copy.deepcopy(grid)
    g = copy.copy(grid)
    for k, v in list(vars(grid).items()):
        base = k.split("_")[0]
        if base in skip and isinstance(v, np.ndarray):
            setattr(g, k, None)
        else:
            setattr(g, k, copy.deepcopy(v))
    return g


def _targets_array(t):
    """Grid.targets (float32 record array x, y, z, r; grid.py:39) -> (n, 4) float32."""
    if t is None or len(t) == 0:
        return np.zeros((0, 4), np.float32)
    t = np.asarray(t)
    if t.dtype.names:
        return np.stack([t["x"], t["y"], t["z"], t["r"]], axis=1).astype(np.float32)
    return t.astype(np.float32).reshape(-1, 4)


def slab_from_env(nx):
    """x-slab of this process under torchrun (RANK / WORLD_SIZE): (x0, nxl, rank, nranks)."""
    rank, n = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    x0, nxl = hm.split_slabs(nx, n)[rank]
    return x0, nxl, rank, n
