"""BASELINE config #1: data/default.json (31x21x6, 6 Au cylinders, sin f=100, courant 0.1, 1000 steps)
through the plugin, surface recording on; the reference's NumPy solver takes ~1.06 s for the same run
(this container)."""
import json, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import helpers as H
from tests.test_gpu_plugin import fake_from_golden
from phonomena_b200.solver_b200 import Solver
d = H.load_golden("default_json_1000")
out = {}
for mode, cfg in (("fp64 exact, surface HDF5", {"arith": "exact", "record": "surface", "write_mode": "thread"}),
                  ("fp64 fast, surface HDF5", {"arith": "fast", "record": "surface", "write_mode": "thread"}),
                  ("fp64 fast, full-field HDF5 (reference schema)", {"arith": "fast", "record": "full", "write_mode": "thread"}),
                  ("fp64 fast, no output", {"arith": "fast", "write_mode": "off"})):
    s = Solver(); s.cfg.update({"wave": "sin", "wave_args": {"f": 100}, "chunk_steps": 250, "kernel": os.environ.get("CFG1_KERNEL", "auto")}); s.cfg.update(cfg)
    s.file = os.path.join(tempfile.gettempdir(), "cfg1.h5")
    g, m = fake_from_golden(d)
    s.init(g, m, 1000); s.run()          # warm
    s.init(g, m, 1000); t = time.perf_counter(); s.run(); dt = time.perf_counter() - t
    ok = all(np.array_equal(a, d[k]) for a, k in zip(s.fields(), ("ux", "uy", "uz"))) if cfg["arith"] == "exact" else H.rel_l2(s.fields(), [d["ux"], d["uy"], d["uz"]])
    out[mode] = {"seconds": round(dt, 4), "mcells_per_s": round(3906 * 1000 / dt / 1e6, 1), "parity": ok if isinstance(ok, bool) else float(ok), "launches": s.stats["launches"]}
print(json.dumps(out, indent=1))
