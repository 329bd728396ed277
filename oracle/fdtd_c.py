"""ctypes wrapper of the C oracle (oracle/fdtd_c.c) -- TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from . import fdtd_numpy as onp

_HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def load(omp=False):
    name = "liboracle_c_omp.so" if omp else "liboracle_c.so"
    if name not in _libs:
        path = os.path.join(_HERE, "_build", name)
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        lib = C.CDLL(path)
        dp, u8 = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        lib.oc_create.restype = C.c_void_p
        lib.oc_create.argtypes = [C.c_int] * 3 + [dp] * 6 + [u8, C.c_int, dp, dp, C.c_double, C.c_double]
        lib.oc_destroy.argtypes = [C.c_void_p]
        lib.oc_array.restype = dp
        lib.oc_array.argtypes = [C.c_void_p, C.c_int]
        lib.oc_step.argtypes = [C.c_void_p, C.c_double]
        lib.oc_run.argtypes = [C.c_void_p, dp, C.c_int]
        _libs[name] = lib
    return _libs[name]


class COracle:
    """Same inputs as the NumPy oracle but material given as (ids, [6x6 tables], [rho])."""
    NAMES = ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old", "T1", "T2", "T3", "T4", "T5", "T6", "ux_new", "uy_new", "uz_new")

    def __init__(self, x, y, z, ids, tables, rhos, dt, wave="sin", wave_args=None, omp=False):
        self.lib = load(omp)
        nx, ny, nz = len(x), len(y), len(z)
        self.shape = (nx, ny, nz)
        fd = [np.ascontiguousarray(a.reshape(-1)) for a in onp.spacings(x, y, z)]
        ids = np.ascontiguousarray(ids, np.uint8)
        assert ids.shape == (nx, ny, nz)
        tab = np.ascontiguousarray(np.array(tables, np.float64).reshape(len(rhos), 36))
        rho = np.ascontiguousarray(np.array(rhos, np.float64))
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        self.dt, self.wave, self.wave_args, self.tt = dt, wave, dict(wave_args or {"f": 100}), 0
        self.h = self.lib.oc_create(nx, ny, nz, *[dp(a) for a in fd], ids.ctypes.data_as(C.POINTER(C.c_uint8)),
                                    len(rho), dp(tab), dp(rho), float(dt), float(dt ** 2))
        shp = {"ux": (nx - 1, ny, nz), "uy": (nx, ny - 1, nz), "uz": (nx, ny, nz - 1), "T1": (nx, ny, nz), "T2": (nx, ny, nz),
               "T3": (nx, ny, nz), "T4": (nx, ny - 1, nz - 1), "T5": (nx - 1, ny, nz - 1), "T6": (nx - 1, ny - 1, nz)}
        for n, name in enumerate(self.NAMES):
            base = name.split("_")[0]
            setattr(self, name, np.ctypeslib.as_array(self.lib.oc_array(self.h, n), shape=shp[base]))

    def step(self, w=None):
        if w is None:
            w = onp.SOURCES[self.wave](tt=self.tt, dt=self.dt, **self.wave_args)
        self.lib.oc_step(self.h, float(w))
        self.tt += 1

    def run(self, steps):
        w = np.ascontiguousarray(onp.source_table(self.wave, self.tt + steps, self.dt, self.wave_args)[self.tt:])
        self.lib.oc_run(self.h, w.ctypes.data_as(C.POINTER(C.c_double)), steps)
        self.tt += steps
        return self

    def close(self):
        if self.h:
            self.lib.oc_destroy(self.h)
            self.h = None
