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


DISCOVER_AND_RUN = '''
import json, os, sys
sys.path.insert(0, %(root)r)
import numpy as np
from oracle import refshim
common = refshim.install()                 # PHONOMENA_REF = a copy of the reference with INTEGRATION.md's solver_b200.py dropped in
common.findSolvers()
assert "b200" in common.solver_dict, list(common.solver_dict)
settings = json.load(open(os.path.join(refshim.REF_ROOT, "data", "default.json")))
settings["simulation"]["solver"] = "b200"
settings["simulation"]["cfg"].update({"write_mode": "thread", "arith": "exact", "record": "auto"})
path = os.path.join(%(tmp)r, "b200.json")
json.dump(settings, open(path, "w"))
cfg, g, m = common.loadSettings(path)
s = common.solver
assert s.name == "b200"
s.file = os.path.join(%(tmp)r, "out.hdf5")
s.init(g, m, 150)
s.run()
r = common.solver_dict["default"]
r.cfg.update({k: v for k, v in settings["simulation"]["cfg"].items() if k in ("wave", "wave_args")})
r.cfg["write_mode"] = "off"
r.init(g, m, 150)
r.run()
for a, b in zip(s.fields(), (r.g.ux, r.g.uy, r.g.uz)):
    assert np.array_equal(a, b)
# the file the GUI would open next (main_widget.py:86-88,129): whole fields per step = the reference's schema, and the last
# frame is the final state
from phonomena_b200 import h5compat
with h5compat.File(s.file, "r") as hdf:
    u = hdf.get("uz")
    assert u.shape == r.g.uz.shape + (150,) and np.array_equal(u[:, :, :, 149], r.g.uz)
    assert np.array_equal(hdf.get("density")[...], r.m.P) and hdf.attrs["dt"] == r.m.dt
print("LIVE PLUGIN OK")
'''


def test_plugin_discovered_and_run_inside_a_reference_checkout(tmp_path):
    """End to end as a maintainer would use it: the three-line solver_b200.py inside a copy of the reference's solvers
    directory, found by common.findSolvers, selected by a settings file, run on the GPU, output file read back."""
    import shutil
    import subprocess
    import sys
    from oracle import refshim
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref = tmp_path / "ref"
    shutil.copytree(os.path.join(refshim.REF_ROOT, "phonomena"), ref / "phonomena", ignore=shutil.ignore_patterns("__pycache__"))
    shutil.copytree(os.path.join(refshim.REF_ROOT, "data"), ref / "data")
    (ref / "phonomena" / "simulation" / "solvers" / "solver_b200.py").write_text(
        "from phonomena_b200.solver_b200 import Solver   # noqa: F401\n"
        "from phonomena_b200.solver_b200 import cfg      # noqa: F401\n")
    script = tmp_path / "run.py"
    script.write_text(DISCOVER_AND_RUN % {"root": root, "tmp": str(tmp_path)})
    env = dict(os.environ, PHONOMENA_REF=str(ref), TMPDIR=str(tmp_path))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert r.returncode == 0 and "LIVE PLUGIN OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
