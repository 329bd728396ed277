"""BASELINE config #5 is a LONG run (10 000 steps).  Tolerances after 10 000 steps, against the bit-identical EXACT
mode (proved equal to the reference on every fixture, tests/test_gpu_parity.py) standing in for the reference, which
would need hours for this on the CPU:

* fp64 FAST stays within the north-star 1e-12;
* plain fp32 does NOT stay within 1e-5 beyond a few thousand steps (the reference code cast to
  float32 drifts the same way: 5.8e-5 at 10 000 steps on default.json) -- asserted <= 2e-4 and reported;
* the compensated fp32 state (arith = "compensated") cuts that several-fold -- asserted better than plain and <= 5e-5;
  what include/phb200.h states about PHB_COMP is this measurement, not "within 1e-5".
Grid: a 96 x 64 x 48 crystal block (the wave crosses it many times in 10 000 steps, absorbing faces at work)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu
STEPS = 10_000


def _run(case, dtype, arith):
    with case.make_engine(steps=STEPS, dtype=dtype, arith=arith) as e:
        e.run(STEPS)
        return e.get_fields()


def test_tolerances_after_10k_steps():
    from phonomena_b200.workloads import crystal_case
    case = crystal_case(96, 64, 48)
    exact = _run(case, "f64", "exact")
    assert all(np.isfinite(a).all() for a in exact) and float(np.abs(exact[2]).max()) > 0.1
    err = {"f64_fast": H.rel_l2(_run(case, "f64", "fast"), exact),
           "f32_plain": H.rel_l2(_run(case, "f32", "fast"), exact),
           "f32_comp": H.rel_l2(_run(case, "f32", "compensated"), exact),
           "f64_comp": H.rel_l2(_run(case, "f64", "compensated"), exact)}
    print("rel-L2 after %d steps vs EXACT: %s" % (STEPS, {k: "%.3e" % v for k, v in err.items()}))
    assert err["f64_fast"] <= 1e-12, err
    assert err["f64_comp"] <= 1e-12, err
    assert err["f32_plain"] <= 2e-4, err
    assert err["f32_comp"] <= 5e-5 and err["f32_comp"] < err["f32_plain"], err


def test_default_json_10k_steps():
    """The reference's own default.json grid for 10 000 steps (stays finite, max |uz| = 1.39)."""
    d = H.load_golden("default_json_1000")
    out = {}
    for key, dtype, arith in (("exact", "f64", "exact"), ("fast", "f64", "fast"), ("f32", "f32", "fast"), ("comp", "f32", "compensated")):
        with H.engine_from_case(d["x"], d["y"], d["z"], d["targets"], d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"], d["courant"],
                                d["wave"], d["wave_args"], STEPS, dtype=dtype, arith=arith) as e:
            e.run(STEPS)
            out[key] = e.get_fields()
    assert abs(float(np.abs(out["exact"][2]).max()) - 1.39) < 0.01
    errs = {k: H.rel_l2(out[k], out["exact"]) for k in ("fast", "f32", "comp")}
    print("default.json, 10 000 steps, rel-L2 vs EXACT:", {k: "%.3e" % v for k, v in errs.items()})
    assert errs["fast"] <= 1e-12 and errs["f32"] <= 2e-4 and errs["comp"] <= 5e-5 and errs["comp"] < errs["f32"]
