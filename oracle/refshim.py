"""Import the UNMODIFIED reference solver in the build container -- TEST INFRASTRUCTURE ONLY.

``/root/reference`` exists only in the build container; ``make -C oracle ref`` stages an
unmodified copy of the solver package in the git-ignored ``oracle/_ref/``, which travels to the GPU
box.  This module is used by ``oracle/gen_golden.py`` (fixture generation), by ``oracle/ref_bench.py``
(the timed CPU baseline of ``bench.py --impl reference``) and by the tests marked ``ref`` (skipped when
neither copy is present).  Nothing in the product imports it.

Three shims, no edits to the reference:
  1. stub ``PyQt5.QtCore`` (``gui/worker.py`` only needs QObject/QRunnable/pyqtSignal/pyqtSlot)
  2. stub ``h5py`` (the solver always runs with ``write_mode='off'``; ``h5py.File.in_memory`` feeds
     simulation/analysis.py with arrays)
  3. ``numpy.float = float`` (removed in NumPy >= 1.24, used at grid.py:258,301)
"""
import os
import sys
import types

import numpy as np

def _find_root():
    """$PHONOMENA_REF, the read-only checkout of the build container, or the copy `make -C oracle ref` staged in
    the git-ignored oracle/_ref/ (which travels to the GPU box)."""
    cands = [os.environ.get("PHONOMENA_REF"), "/root/reference", os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "phonomena", "simulation")):
            return c
    return cands[1]


REF_ROOT = _find_root()


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "phonomena", "simulation"))


class _Signal:
    def __init__(self, *a, **k):
        pass

    def emit(self, *a, **k):
        pass

    def connect(self, *a, **k):
        pass


def install():
    """Make ``import common`` / ``from simulation import ...`` resolve to the reference."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REF_ROOT)
    if not hasattr(np, "float"):
        np.float = float
    if "PyQt5" not in sys.modules:
        qt, qc = types.ModuleType("PyQt5"), types.ModuleType("PyQt5.QtCore")

        class QObject:
            def __init__(self, *a, **k):
                pass

        class QRunnable(QObject):
            pass

        qc.QObject, qc.QRunnable = QObject, QRunnable
        qc.pyqtSignal = lambda *a, **k: _Signal()
        qc.pyqtSlot = lambda *a, **k: (lambda f: f)
        qt.QtCore = qc
        sys.modules["PyQt5"], sys.modules["PyQt5.QtCore"] = qt, qc
    if "h5py" not in sys.modules:
        h5 = types.ModuleType("h5py")

        class _Dataset:
            """What simulation/analysis.py touches of an h5py dataset: != None, .shape, slicing."""

            def __init__(self, a):
                self._a = np.asarray(a)
                self.shape = self._a.shape

            def __ne__(self, other):
                return True

            def __getitem__(self, idx):
                return self._a[idx]

        class File:
            """In-memory stand-in: real files are refused (the reference solver runs with
            write_mode='off'); ``File.in_memory(datasets, attrs)`` feeds simulation/analysis.py."""

            def __init__(self, *a, **k):
                raise RuntimeError("h5py is stubbed: run the reference with write_mode='off'")

            @classmethod
            def in_memory(cls, datasets, attrs):
                f = cls.__new__(cls)
                f._d = {k: _Dataset(v) for k, v in datasets.items()}
                f.attrs = dict(attrs)
                return f

            def get(self, name):
                return self._d.get(name)

            def close(self):
                pass

        h5.File = File
        h5.__stub__ = True
        sys.modules["h5py"] = h5
    pkg = os.path.join(REF_ROOT, "phonomena")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    import common  # noqa: F401  (reference module)
    return common


def default_solver():
    """A fresh reference ``solver_default.Solver`` with the writer off."""
    install()
    from simulation.solvers import solver_default
    s = solver_default.Solver()
    s.cfg["write_mode"] = "off"
    return s
