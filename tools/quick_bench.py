"""Ad-hoc kernel timing (development aid; bench.py is the contract)."""
import argparse, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from phonomena_b200 import _lib, hostmath as hm
from phonomena_b200.workloads import crystal_case

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, nargs="+", default=[256, 256, 256])
ap.add_argument("--dtype", default="f64")
ap.add_argument("--arith", default="fast")
ap.add_argument("--kernel", default="auto")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--homog", action="store_true")
a = ap.parse_args()
nx, ny, nz = a.n
case = crystal_case(nx, ny, nz, homogeneous=a.homog)
e = case.make_engine(dtype=a.dtype, arith=a.arith, kernel=a.kernel, steps=a.steps + a.warmup)
e.run(a.warmup); e.sync()
ms = e.run_timed(a.steps)
cells = nx * ny * nz
es = 8 if a.dtype == "f64" else 4
try:      # the driver-written measured copy bandwidth (bench.py uses the same file), else the profiling guide's fallback
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
balg = 9 * es + 1
print(json.dumps({"n": a.n, "dtype": a.dtype, "arith": a.arith, "kernel": e.info()["kernel"], "ms_per_step": ms / a.steps,
                  "gcells": cells * a.steps / ms / 1e6, "GBs_alg": cells * balg * a.steps / ms / 1e6,
                  "frac_of_hbm_peak": cells * balg * a.steps / ms / 1e6 / PEAK, "launches": e.launch_count}))
