import sys; sys.path.insert(0,'.')
import numpy as np
from tests import helpers as H
d = H.load_golden("default_json_1000")
for steps in (1000, 10000):
    d["steps"]=steps
    out={}
    for key,dt,ar in (("exact","f64","exact"),("fast","f64","fast"),("f32","f32","fast"),("f32c","f32","comp"),("f64c","f64","comp")):
        with H.engine_from_golden(d, dtype=dt, arith=ar) as e:
            e.run(steps); out[key]=e.get_fields()
    print(steps, "fast %.2e f32 %.2e f32-comp %.2e f64-comp %.2e" % tuple(H.rel_l2(out[k],out["exact"]) for k in ("fast","f32","f32c","f64c")))
