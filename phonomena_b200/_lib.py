"""ctypes binding of libphb200.so (C ABI in include/phb200.h).

There is no CPU fallback: if the shared library is missing or no B200 is visible the
calls raise.  PyTorch is not needed here; NumPy arrays are the host buffers.
"""
from __future__ import annotations

import atexit
import ctypes as C
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# PHB200_LIB: development override (timing variants built by tools/diag_build.sh); still a CUDA library, not a fallback
LIB_PATH = os.environ.get("PHB200_LIB") or os.path.join(_HERE, "libphb200.so")

F32, F64 = 0, 1
FAST, EXACT, COMP = 0, 1, 2
KERNEL_AUTO, KERNEL_NAIVE, KERNEL_MARCH = 0, 1, 2
CUR, OLD = 0, 1
REC_UX, REC_UY, REC_UZ, REC_FULL = 1, 2, 4, 8
BC_ABSORBING, BC_PERIODIC = 0, 1
WRITER_MMAP, WRITER_POPULATE = 1, 2

# every symbol include/phb200.h declares (tests check the .so exports all of them)
SYMBOLS = (
    "phb_version", "phb_last_error", "phb_device_count", "phb_create", "phb_destroy",
    "phb_set_spacing", "phb_set_material_table", "phb_set_material_dense", "phb_set_material_ids", "phb_gen_material_ids",
    "phb_get_material_ids", "phb_set_abc", "phb_set_source_table", "phb_set_fields", "phb_get_fields",
    "phb_get_stress", "phb_run", "phb_sync", "phb_run_timed", "phb_steps_done", "phb_launch_count",
    "phb_info", "phb_profile", "phb_comm_unique_id", "phb_comm_init", "phb_p2p_export", "phb_p2p_import", "phb_p2p_mode", "phb_bloch_pair", "phb_record_next", "phb_record_release",
    "phb_record_frame_doubles", "phb_record_abort", "phb_record_timeout", "phb_cancel", "phb_writer_start", "phb_writer_mapped", "phb_writer_finish",
    "phb_writer_selftest", "phb_probe_add", "phb_probe_shape", "phb_probe_read", "phb_probe_dft_t", "phb_probe_dft_xt",
)


class PhbError(RuntimeError):
    pass


class PhbCancelled(PhbError):
    """phb_run returned because phb_cancel was called (not an error for the plugin's run loop)."""


class Cfg(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("x0", C.c_int32), ("nxl", C.c_int32),
        ("dtype", C.c_int32), ("arith", C.c_int32), ("device", C.c_int32), ("kernel", C.c_int32),
        ("record_mask", C.c_int32), ("record_every", C.c_int32), ("ring_slots", C.c_int32),
        ("bc_y", C.c_int32), ("record_stride", C.c_int32 * 3),
        ("dt", C.c_double), ("d2", C.c_double),
    ]


_lib = None


def load_library(path=None):
    """Load libphb200.so; raises PhbError if it has not been built (run `python -c
    "import __graft_entry__ as g; g.build()"` or `make -C phonomena_b200/csrc`)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise PhbError("libphb200.so not found at %s: build it first (make -C phonomena_b200/csrc); "
                       "there is no CPU fallback" % p)
    lib = C.CDLL(p)
    dp, u8p, fp, vp = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.c_void_p
    lib.phb_version.restype = C.c_int
    lib.phb_last_error.restype = C.c_char_p
    lib.phb_device_count.argtypes = [C.POINTER(C.c_int)]
    lib.phb_create.argtypes = [C.POINTER(Cfg), C.POINTER(vp)]
    lib.phb_destroy.argtypes = [vp]
    lib.phb_set_spacing.argtypes = [vp] + [dp] * 6
    lib.phb_set_material_table.argtypes = [vp, C.c_int32, dp, dp]
    lib.phb_set_material_ids.argtypes = [vp, u8p, C.c_int64]
    lib.phb_set_material_dense.argtypes = [vp, dp, dp, C.c_int64]
    lib.phb_gen_material_ids.argtypes = [vp, fp, C.c_int32, dp, dp, dp]
    lib.phb_get_material_ids.argtypes = [vp, u8p]
    lib.phb_set_abc.argtypes = [vp, dp]
    lib.phb_set_source_table.argtypes = [vp, dp, C.c_int64]
    lib.phb_set_fields.argtypes = [vp, C.c_int32, dp, dp, dp]
    lib.phb_get_fields.argtypes = [vp, C.c_int32, dp, dp, dp]
    lib.phb_get_stress.argtypes = [vp, C.c_int32] + [dp] * 6
    lib.phb_run.argtypes = [vp, C.c_int64]
    lib.phb_sync.argtypes = [vp]
    lib.phb_run_timed.argtypes = [vp, C.c_int64, C.POINTER(C.c_float)]
    lib.phb_steps_done.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.phb_launch_count.argtypes = [vp, C.POINTER(C.c_int64)]
    lib.phb_info.argtypes = [vp, C.c_char_p, C.c_int32, C.POINTER(C.c_int64)]
    lib.phb_profile.argtypes = [vp, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    lib.phb_comm_unique_id.argtypes = [C.c_char_p]
    lib.phb_comm_init.argtypes = [vp, C.c_char_p, C.c_int32, C.c_int32]
    lib.phb_p2p_export.argtypes = [vp, C.c_char_p, C.POINTER(C.c_int32)]
    lib.phb_p2p_import.argtypes = [vp, C.c_int32, C.c_int32, C.c_char_p, C.c_int32, C.c_char_p, C.c_int32]
    lib.phb_p2p_mode.argtypes = [vp, C.c_int32]
    lib.phb_bloch_pair.argtypes = [vp, vp, C.c_double]
    lib.phb_record_next.argtypes = [vp, C.POINTER(dp), C.POINTER(C.c_int64), C.c_int32]
    lib.phb_record_release.argtypes = [vp]
    lib.phb_record_frame_doubles.argtypes = [vp, C.POINTER(C.c_int64)]
    i64p = C.POINTER(C.c_int64)
    lib.phb_record_abort.argtypes = [vp, C.c_char_p]
    lib.phb_record_timeout.argtypes = [vp, C.c_int32]
    lib.phb_cancel.argtypes = [vp]
    lib.phb_writer_start.argtypes = [vp, C.c_int32, C.c_int32, i64p, i64p, C.c_int64, C.c_int64, C.c_int32, C.c_int32]
    lib.phb_writer_mapped.argtypes = [vp, C.POINTER(C.c_int32)]
    lib.phb_writer_finish.argtypes = [vp, C.c_int32, i64p, dp, dp]
    lib.phb_writer_selftest.argtypes = [C.c_int32, C.c_int32, i64p, i64p, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, i64p]
    lib.phb_probe_add.argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.POINTER(C.c_int32)]
    lib.phb_probe_shape.argtypes = [vp, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.phb_probe_read.argtypes = [vp, C.c_int32, dp]
    lib.phb_probe_dft_t.argtypes = [vp, C.c_int32, dp, C.c_int64, C.c_int64, C.c_int64, dp, dp]
    lib.phb_probe_dft_xt.argtypes = [vp, C.c_int32, dp, C.c_int64, C.c_int64, dp, dp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("phb_last_error",):
            fn.restype = C.c_int
    if path is None:
        _lib = lib
    return lib


def _chk(lib, rc):
    if rc != 0:
        raise PhbError(lib.phb_last_error().decode("utf-8", "replace"))


def _dptr(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise PhbError("array has shape %s, expected %s" % (a.shape, tuple(shape)))
    return a


def device_count():
    lib = load_library()
    n = C.c_int(0)
    lib.phb_device_count(C.byref(n))
    return n.value


def writer_selftest(fd, base, nbytes, stride, frames, slots=4, nthreads=2, timeout_ms=2000, mmap=False):
    """Host-only run of the recorder ring + native writer threads (no CUDA call); returns frames written."""
    lib = load_library()
    n = len(base)
    b = (C.c_int64 * n)(*[int(v) for v in base])
    nb = (C.c_int64 * n)(*[int(v) for v in nbytes])
    w = C.c_int64(0)
    rc = lib.phb_writer_selftest(int(fd), n, b, nb, int(stride), int(frames), int(slots), int(nthreads), int(timeout_ms),
                                 WRITER_MMAP if mmap else 0, C.byref(w))
    if rc != 0:
        raise PhbError("%s (after %d frames)" % (lib.phb_last_error().decode("utf-8", "replace"), w.value))
    return w.value


def comm_unique_id():
    lib = load_library()
    buf = C.create_string_buffer(128)
    _chk(lib, lib.phb_comm_unique_id(buf))
    return buf.raw


_live = weakref.WeakSet()


def _close_all():
    """Interpreter exit: destroy the contexts that are still alive NOW, while the CUDA runtime still is -- a context
    destroyed from __del__ during interpreter teardown (after libcudart's own exit handlers) can crash the process."""
    for e in list(_live):
        try:
            e.close()
        except Exception:
            pass


atexit.register(_close_all)


class Engine:
    """One device context = one grid or one x-slab [x0, x0+nxl) on one GPU."""

    def __init__(self, nx, ny, nz, dt, d2=None, dtype="f64", arith="fast", device=0, x0=0, nxl=None,
                 kernel="auto", record_mask=0, record_every=1, ring_slots=0, bc_y="absorbing", record_stride=(1, 1, 1)):
        self.lib = load_library()
        self.nx, self.ny, self.nz = int(nx), int(ny), int(nz)
        self.x0 = int(x0)
        self.nxl = int(self.nx - self.x0 if nxl is None else nxl)
        self.dtype = {"f32": F32, "fp32": F32, "f64": F64, "fp64": F64}[dtype]
        self.arith = {"fast": FAST, "exact": EXACT, "compensated": COMP, "comp": COMP}[arith]
        cfg = Cfg()
        cfg.nx, cfg.ny, cfg.nz, cfg.x0, cfg.nxl = self.nx, self.ny, self.nz, self.x0, self.nxl
        cfg.dtype, cfg.arith, cfg.device = self.dtype, self.arith, int(device)
        cfg.kernel = {"auto": KERNEL_AUTO, "naive": KERNEL_NAIVE, "march": KERNEL_MARCH}[kernel]
        cfg.record_mask, cfg.record_every, cfg.ring_slots = int(record_mask), int(record_every), int(ring_slots)
        self.record_stride = tuple(max(1, int(v)) for v in record_stride) if (int(record_mask) & REC_FULL) else (1, 1, 1)
        for q in range(3):
            cfg.record_stride[q] = self.record_stride[q]
        cfg.bc_y = {"absorbing": BC_ABSORBING, "periodic": BC_PERIODIC}[bc_y]
        self.bc_y = bc_y
        cfg.dt = float(dt)
        # the reference evaluates self.m.dt**2 in Python (base_solver.py:443)
        cfg.d2 = float(dt ** 2 if d2 is None else d2)
        self.record_mask = int(record_mask)
        self._ctx = C.c_void_p()
        _chk(self.lib, self.lib.phb_create(C.byref(cfg), C.byref(self._ctx)))
        _live.add(self)

    # -- lifetime ---------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self.lib.phb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- shapes -------------------------------------------------------------------------
    def planes(self, comp):
        hi = self.nx - 1 if comp == 0 else self.nx
        return max(0, min(self.x0 + self.nxl, hi) - self.x0)

    def shapes(self):
        ny, nz = self.ny, self.nz
        return ((self.planes(0), ny, nz), (self.planes(1), ny - 1, nz), (self.planes(2), ny, nz - 1))

    # -- setup --------------------------------------------------------------------------
    def set_spacing(self, fdx, fdy, fdz, sdx, sdy, sdz):
        n = (self.nx - 1, self.ny - 1, self.nz - 1, self.nx - 2, self.ny - 2, self.nz - 2)
        arrs = [_f64(np.asarray(a).reshape(-1), (m,)) for a, m in zip((fdx, fdy, fdz, sdx, sdy, sdz), n)]
        _chk(self.lib, self.lib.phb_set_spacing(self._ctx, *[_dptr(a) for a in arrs]))

    def set_material_table(self, tables, rhos):
        """tables: list of 6x6 (already scaled) stiffness matrices; rhos: densities."""
        c12 = np.array([[t[r][c] for r in range(3) for c in range(3)] + [t[3][3], t[4][4], t[5][5]]
                        for t in (np.asarray(t, np.float64) for t in tables)], np.float64)
        rho = _f64(np.asarray(rhos, np.float64).reshape(-1), (len(c12),))
        c12 = _f64(c12, (len(rho), 12))
        _chk(self.lib, self.lib.phb_set_material_table(self._ctx, len(rho), _dptr(c12), _dptr(rho)))

    def id_planes(self):
        return min(self.x0 + self.nxl + 1, self.nx) - self.x0

    def set_material_ids(self, ids):
        ids = np.ascontiguousarray(ids, dtype=np.uint8)
        if ids.shape != (self.id_planes(), self.ny, self.nz):
            raise PhbError("ids has shape %s, expected %s" % (ids.shape, (self.id_planes(), self.ny, self.nz)))
        _chk(self.lib, self.lib.phb_set_material_ids(self._ctx, ids.ctypes.data_as(C.POINTER(C.c_uint8)), ids.shape[0]))

    def set_material_dense(self, Cd, Pd):
        """The reference's own arrays: C (planes, Ny, Nz, 6, 6), P (planes, Ny, Nz) for planes
        x0 .. x0 + id_planes() - 1 (Material.C / Material.P, material.py:48-63)."""
        n = self.id_planes()
        Cd, Pd = _f64(Cd, (n, self.ny, self.nz, 6, 6)), _f64(Pd, (n, self.ny, self.nz))
        _chk(self.lib, self.lib.phb_set_material_dense(self._ctx, _dptr(Cd), _dptr(Pd), n))

    def gen_material_ids(self, targets, x, y, z):
        """targets: (n,4) float32 rows x, y, z, r -- evaluated on the device."""
        t = np.ascontiguousarray(np.asarray(targets, np.float32).reshape(-1, 4))
        x, y, z = _f64(x, (self.nx,)), _f64(y, (self.ny,)), _f64(z, (self.nz,))
        _chk(self.lib, self.lib.phb_gen_material_ids(self._ctx, t.ctypes.data_as(C.POINTER(C.c_float)), len(t),
                                                     _dptr(x), _dptr(y), _dptr(z)))

    def get_material_ids(self):
        out = np.empty((self.nxl, self.ny, self.nz), np.uint8)
        _chk(self.lib, self.lib.phb_get_material_ids(self._ctx, out.ctypes.data_as(C.POINTER(C.c_uint8))))
        return out

    def set_abc(self, coef):
        """coef: dict or sequence clx, ctx, cly0, cty0, cly1, cty1, clz, ctz."""
        if isinstance(coef, dict):
            coef = [coef[k] for k in ("clx", "ctx", "cly0", "cty0", "cly1", "cty1", "clz", "ctz")]
        a = _f64(np.array([float(v) for v in coef]), (8,))
        _chk(self.lib, self.lib.phb_set_abc(self._ctx, _dptr(a)))

    def set_source_table(self, w):
        if w is None:
            _chk(self.lib, self.lib.phb_set_source_table(self._ctx, None, 0))
            return
        w = _f64(np.asarray(w).reshape(-1))
        _chk(self.lib, self.lib.phb_set_source_table(self._ctx, _dptr(w), len(w)))

    # -- fields -------------------------------------------------------------------------
    def set_fields(self, ux=None, uy=None, uz=None, which=CUR):
        sh = self.shapes()
        arrs = [None if a is None else _f64(a, s) for a, s in zip((ux, uy, uz), sh)]
        _chk(self.lib, self.lib.phb_set_fields(self._ctx, which, *[_dptr(a) for a in arrs]))

    def get_fields(self, which=CUR):
        out = [np.empty(s, np.float64) for s in self.shapes()]
        _chk(self.lib, self.lib.phb_get_fields(self._ctx, which, *[_dptr(a) for a in out]))
        return out

    def get_stress(self, which=OLD):
        n, n5, ny, nz = self.nxl, self.planes(0), self.ny, self.nz
        shp = [(n, ny, nz)] * 3 + [(n, ny - 1, nz - 1), (n5, ny, nz - 1), (n5, ny - 1, nz)]
        out = [np.zeros(s, np.float64) for s in shp]
        _chk(self.lib, self.lib.phb_get_stress(self._ctx, which, *[_dptr(a) for a in out]))
        return out

    # -- stepping -----------------------------------------------------------------------
    def run(self, nsteps):
        rc = self.lib.phb_run(self._ctx, int(nsteps))
        if rc == 3:
            raise PhbCancelled("cancelled")
        _chk(self.lib, rc)

    def cancel(self):
        """Make the phb_run in progress (on another thread) return after the current step."""
        if self._ctx:
            self.lib.phb_cancel(self._ctx)

    def sync(self):
        _chk(self.lib, self.lib.phb_sync(self._ctx))

    def run_timed(self, nsteps):
        ms = C.c_float(0)
        _chk(self.lib, self.lib.phb_run_timed(self._ctx, int(nsteps), C.byref(ms)))
        return ms.value

    @property
    def steps_done(self):
        n = C.c_int64(0)
        _chk(self.lib, self.lib.phb_steps_done(self._ctx, C.byref(n)))
        return n.value

    @property
    def launch_count(self):
        n = C.c_int64(0)
        _chk(self.lib, self.lib.phb_launch_count(self._ctx, C.byref(n)))
        return n.value

    def info(self):
        name = C.create_string_buffer(64)
        nbytes = C.c_int64(0)
        _chk(self.lib, self.lib.phb_info(self._ctx, name, 64, C.byref(nbytes)))
        return {"kernel": name.value.decode(), "device_bytes": nbytes.value}

    def profile(self, enable=-1):
        """(ms, launches) spent in the fused stencil kernel since profiling was (re)enabled."""
        ms, n = C.c_double(0), C.c_int64(0)
        _chk(self.lib, self.lib.phb_profile(self._ctx, int(enable), C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def bloch_pair(self, imag, phase):
        """Make (self, imag) the real / imaginary part of a Bloch-periodic field with u(y + L) = u(y) exp(i phase);
        run() on self then steps both (include/phb200.h phb_bloch_pair)."""
        _chk(self.lib, self.lib.phb_bloch_pair(self._ctx, imag._ctx, float(phase)))
        self._bloch_imag = imag      # keep the partner alive as long as the primary

    # -- multi-GPU ----------------------------------------------------------------------
    def comm_init(self, unique_id, rank, nranks):
        _chk(self.lib, self.lib.phb_comm_init(self._ctx, unique_id, int(rank), int(nranks)))

    def p2p_export(self):
        """(256-byte IPC handle blob, nxl) for the fused NVLink halo push."""
        buf, nxl = C.create_string_buffer(256), C.c_int32(0)
        _chk(self.lib, self.lib.phb_p2p_export(self._ctx, buf, C.byref(nxl)))
        return buf.raw, nxl.value

    def p2p_import(self, rank, nranks, exports):
        """exports: list over ranks of p2p_export() results."""
        left = exports[rank - 1] if rank > 0 else (None, 0)
        right = exports[rank + 1] if rank < nranks - 1 else (None, 0)
        _chk(self.lib, self.lib.phb_p2p_import(self._ctx, int(rank), int(nranks), left[0], int(left[1]), right[0], int(right[1])))

    def connect(self, rank, nranks, allgather, broadcast, mode=None):
        """Set up the halo exchange of an x-slab context.  mode 'p2p' (default: a push kernel stores the finished edge
        planes into the neighbours' ghost planes over NVLink, CUDA IPC), 'fused' (the stencil kernel stores them itself)
        or 'nccl'; `allgather(obj) -> list` and `broadcast(obj) -> obj` are host channels."""
        import os
        mode = mode or os.environ.get("PHB_HALO", "p2p")
        if nranks == 1:
            return "none"
        if mode in ("p2p", "fused"):
            # every rank runs the same two collectives whatever fails locally (a rank that skipped one would pair
            # its next all-gather with the others' previous one)
            try:
                mine = self.p2p_export()
            except PhbError:
                mine = None
            exports = allgather(mine)
            ok = int(all(x is not None for x in exports))
            if ok:
                try:
                    self.p2p_import(rank, nranks, exports)
                except PhbError:
                    ok = 0
            if all(allgather(ok)):
                _chk(self.lib, self.lib.phb_p2p_mode(self._ctx, 1 if mode == "fused" else 0))
                return mode
        self.comm_init(broadcast(comm_unique_id() if rank == 0 else None), rank, nranks)
        return "nccl"

    # -- recorder -----------------------------------------------------------------------
    def frame_doubles(self):
        n = C.c_int64(0)
        _chk(self.lib, self.lib.phb_record_frame_doubles(self._ctx, C.byref(n)))
        return n.value

    def frame_layout(self):
        """[(name, shape)] of the recorded components inside one frame."""
        out = []
        full = bool(self.record_mask & REC_FULL)      # whole arrays (every record_stride-th entry) instead of their k = 0 planes
        sx, sy, sz = self.record_stride
        cd = lambda n, s: -(-n // s)
        if self.record_mask & REC_UX:
            out.append(("ux", (cd(self.planes(0), sx), cd(self.ny, sy), cd(self.nz, sz)) if full else (self.planes(0), self.ny)))
        if self.record_mask & REC_UY:
            out.append(("uy", (cd(self.nxl, sx), cd(self.ny - 1, sy), cd(self.nz, sz)) if full else (self.nxl, self.ny - 1)))
        if self.record_mask & REC_UZ:
            out.append(("uz", (cd(self.nxl, sx), cd(self.ny, sy), cd(self.nz - 1, sz)) if full else (self.nxl, self.ny)))
        return out

    def record_next(self, timeout_ms=1000):
        """Next recorded frame as (tt, {name: array view into the pinned ring}) or None on timeout.
        The views are valid until record_release()."""
        p = C.POINTER(C.c_double)()
        tt = C.c_int64(0)
        rc = self.lib.phb_record_next(self._ctx, C.byref(p), C.byref(tt), int(timeout_ms))
        if rc == 2:
            return None
        _chk(self.lib, rc)
        flat = np.ctypeslib.as_array(p, shape=(self.frame_doubles(),))
        views, off = {}, 0
        for name, shp in self.frame_layout():
            n = int(np.prod(shp))
            views[name] = flat[off:off + n].reshape(shp)
            off += n
        return tt.value, views

    def record_release(self):
        _chk(self.lib, self.lib.phb_record_release(self._ctx))

    def record_abort(self, why="aborted by the consumer"):
        if self._ctx:
            self.lib.phb_record_abort(self._ctx, str(why).encode("utf-8", "replace")[:500])

    def record_timeout(self, timeout_ms):
        _chk(self.lib, self.lib.phb_record_timeout(self._ctx, int(timeout_ms)))

    def writer_start(self, fd, base, nbytes, stride, frames, nthreads=4, mmap=False, populate=False):
        """Native writer threads: component c of recorded frame f goes to file offset base[c] + f * stride.
        mmap: the extents are allocated and `fd` is read-write -> map them and copy in parallel (else pwrite)."""
        n = len(base)
        b = (C.c_int64 * n)(*[int(v) for v in base])
        nb = (C.c_int64 * n)(*[int(v) for v in nbytes])
        flags = (WRITER_MMAP if mmap else 0) | (WRITER_POPULATE if (mmap and populate) else 0)
        _chk(self.lib, self.lib.phb_writer_start(self._ctx, int(fd), n, b, nb, int(stride), int(frames), int(nthreads), flags))

    def writer_mapped(self):
        m = C.c_int32(0)
        _chk(self.lib, self.lib.phb_writer_mapped(self._ctx, C.byref(m)))
        return bool(m.value)

    def writer_finish(self, timeout_ms=300000):
        """Drain + join; returns (frames written, seconds waiting for frames, seconds writing).  Raises on a
        write error or timeout; the counters are also left in self.writer_stats."""
        n, wa, wr = C.c_int64(0), C.c_double(0), C.c_double(0)
        rc = self.lib.phb_writer_finish(self._ctx, int(timeout_ms), C.byref(n), C.byref(wa), C.byref(wr))
        self.writer_stats = (n.value, wa.value, wr.value)
        _chk(self.lib, rc)
        return self.writer_stats

    # ---- line probes and on-device spectra (include/phb200.h "line probes") ----
    def probe_add(self, comp, j, k, capacity):
        """Keep u_comp[:, j, k] of every recorded step on the device; comp: 'ux'|'uy'|'uz' or 0..2."""
        comp = {"ux": 0, "uy": 1, "uz": 2}.get(comp, comp)
        pid = C.c_int32(-1)
        _chk(self.lib, self.lib.phb_probe_add(self._ctx, int(comp), int(j), int(k), int(capacity), C.byref(pid)))
        return pid.value

    def probe_shape(self, pid):
        rows, frames = C.c_int64(0), C.c_int64(0)
        _chk(self.lib, self.lib.phb_probe_shape(self._ctx, int(pid), C.byref(rows), C.byref(frames)))
        return rows.value, frames.value

    def probe_read(self, pid):
        """(rows, frames) float64: this slab's part of u[:, j, k, :]."""
        rows, frames = self.probe_shape(pid)
        out = np.zeros((rows, frames), np.float64)
        _chk(self.lib, self.lib.phb_probe_read(self._ctx, int(pid), _dptr(out)))
        return out

    def probe_dft_t(self, pid, window, nf, row0=0, nrows=None):
        """Unnormalised sum_t u[row, t] w[t] exp(-2 pi i k t / N) for k < nf: complex (nrows, nf)."""
        rows, frames = self.probe_shape(pid)
        nrows = rows - row0 if nrows is None else nrows
        window = np.ascontiguousarray(window, np.float64)
        if window.size != frames:
            raise ValueError("window has %d weights, the probe holds %d frames" % (window.size, frames))
        re, im = np.zeros((nrows, nf)), np.zeros((nrows, nf))
        _chk(self.lib, self.lib.phb_probe_dft_t(self._ctx, int(pid), _dptr(window), int(nf), int(row0), int(nrows), _dptr(re), _dptr(im)))
        return re + 1j * im

    def probe_dft_xt(self, pid, window, nf, nx_total):
        """This slab's share of the unnormalised 2-D (x, t) transform: complex (nx_total, nf); slabs add."""
        _, frames = self.probe_shape(pid)
        window = np.ascontiguousarray(window, np.float64)
        if window.size != frames:
            raise ValueError("window has %d weights, the probe holds %d frames" % (window.size, frames))
        re, im = np.zeros((nx_total, nf)), np.zeros((nx_total, nf))
        _chk(self.lib, self.lib.phb_probe_dft_xt(self._ctx, int(pid), _dptr(window), int(nf), int(nx_total), _dptr(re), _dptr(im)))
        return re + 1j * im
