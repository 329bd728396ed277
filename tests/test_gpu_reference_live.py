"""The drop-in claim with the LIVE reference on the GPU box: `make -C oracle ref` stages the unmodified reference in
oracle/_ref/, which travels with the snapshot, so here the reference's own loader (common.loadSettings), its own Grid /
Material objects and its own solver_default run side by side with the B200 plugin on the same objects -- no fixture, no
restatement in between.  fp64 EXACT: every entry of ux, uy, uz bit-identical; FAST <= 1e-12; fp32 <= 1e-5."""
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = [pytest.mark.gpu, pytest.mark.ref]


def _reference_run(common, refshim, g, m, scfg, steps):
    r = refshim.default_solver()
    r.cfg.update(scfg)
    r.cfg["write_mode"] = "off"
    r.init(g, m, steps)
    r.run()
    return r


@pytest.mark.parametrize("settings,steps", [("data/default.json", 300), ("tests/data/nonuniform.json", 120)])
def test_plugin_equals_live_reference_on_its_own_settings_files(settings, steps, tmp_path):
    from oracle import refshim
    from phonomena_b200.solver_b200 import Solver
    common = refshim.install()
    common.findSolvers()
    cfg, g, m = common.loadSettings(os.path.join(refshim.REF_ROOT, settings))
    scfg = dict(cfg["simulation"]["cfg"])
    scfg.pop("write_mode", None)
    ref = _reference_run(common, refshim, g, m, scfg, steps)
    want = [ref.g.ux, ref.g.uy, ref.g.uz]
    assert float(np.abs(want[2]).max()) > 0
    x_before = g.x.copy()
    for precision, arith, tol in (("fp64", "exact", 0.0), ("fp64", "fast", 1e-12), ("fp32", "fast", 1e-5)):
        s = Solver()
        s.cfg.update(scfg)
        s.cfg.update({"write_mode": "off", "precision": precision, "arith": arith})
        s.file = str(tmp_path / "live.h5")
        s.init(g, m, steps)                      # the reference's own Grid / Material objects
        assert s.dt == ref.m.dt
        s.run()
        got = s.fields()
        if tol == 0.0:
            for a, b, n in zip(got, want, "xyz"):
                assert np.array_equal(a, b), (settings, "u" + n, float(np.abs(a - b).max()))
        else:
            assert H.rel_l2(got, want) <= tol, (settings, precision, arith)
    assert np.array_equal(g.x, x_before)         # the caller's objects are untouched


def test_solver_test_on_reference_testdefaults():
    """Solver.test() (base_solver.py:286-292) on the reference's real TestDefaults objects, against the reference's own
    test()."""
    from oracle import refshim
    from phonomena_b200.solver_b200 import Solver
    refshim.install()
    r = refshim.default_solver()
    r.test()
    s = Solver()
    s.cfg.update({"write_mode": "off", "arith": "exact", "wave": r.cfg["wave"], "wave_args": dict(r.cfg["wave_args"])})
    s.test()
    for a, b in zip(s.fields(), (r.g.ux, r.g.uy, r.g.uz)):
        assert np.array_equal(a, b)
    assert s.stats["steps"] == 10
