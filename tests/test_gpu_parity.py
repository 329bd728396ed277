"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

fp64 EXACT arithmetic: bit-for-bit (np.array_equal) on every entry of every array,
boundaries included.  fp64 FAST: rel-L2 <= 1e-12.  fp32: rel-L2 <= 1e-5 (BASELINE.json)."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

KERNELS = ["naive", "march"]
TOL_F64_FAST = 1e-12
TOL_F32 = 1e-5


@pytest.mark.parametrize("name", H.golden_names())
def test_device_material_indexing_bit_exact(name):
    d = H.load_golden(name)
    with H.engine_from_golden(d) as e:
        assert np.array_equal(e.get_material_ids(), d["ids"])
        assert e.dt == d["dt"]


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", H.golden_names())
def test_fp64_exact_bitwise_vs_reference(name, kernel):
    d = H.load_golden(name)
    snaps = sorted(int(s) for s in d["snap_steps"])
    with H.engine_from_golden(d, dtype="f64", arith="exact", kernel=kernel) as e:
        done = 0
        for n in snaps:
            e.run(n - done)
            done = n
            ux, uy, uz = e.get_fields()
            assert np.array_equal(ux[:, :, 0], d["snap_ux_%d" % n]), (name, n)
            assert np.array_equal(uy[:, :, 0], d["snap_uy_%d" % n]), (name, n)
            assert np.array_equal(uz[:, :, 0], d["snap_uz_%d" % n]), (name, n)
        e.run(d["steps"] - done)
        assert e.steps_done == d["steps"]
        for got, key in zip(e.get_fields(), ("ux", "uy", "uz")):
            assert np.array_equal(got, d[key]), (name, key, float(np.abs(got - d[key]).max()))
        for got, key in zip(e.get_fields(which=1), ("ux_old", "uy_old", "uz_old")):
            assert np.array_equal(got, d[key]), (name, key)
        for got, key in zip(e.get_stress(which=1), ("T1", "T2", "T3", "T4", "T5", "T6")):
            assert np.array_equal(got, d[key]), (name, key)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", H.golden_names())
def test_fp64_fast_within_1e12(name, kernel):
    d = H.load_golden(name)
    with H.engine_from_golden(d, dtype="f64", arith="fast", kernel=kernel) as e:
        e.run(d["steps"])
        got = e.get_fields()
    err = H.rel_l2(got, [d["ux"], d["uy"], d["uz"]])
    assert err <= TOL_F64_FAST, (name, err)


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", H.golden_names())
def test_fp32_within_1e5(name, kernel):
    d = H.load_golden(name)
    with H.engine_from_golden(d, dtype="f32", arith="fast", kernel=kernel) as e:
        e.run(d["steps"])
        got = e.get_fields()
    err = H.rel_l2(got, [d["ux"], d["uy"], d["uz"]])
    assert err <= TOL_F32, (name, err)


@pytest.mark.parametrize("kernel", KERNELS)
def test_random_fields_nonuniform_xyz_vs_oracle(kernel):
    """Random initial u / u_old on a mesh non-uniform in x, y and z, a few steps, every
    entry compared bit for bit with the oracle (the reference-vs-oracle twin of this test
    is tests/test_oracle_golden.py::test_oracle_vs_live_reference_random_fields)."""
    from oracle import fdtd_numpy as onp
    d = H.load_golden("nonuniform_json_200")
    rng = np.random.default_rng(0)
    z = np.cumsum(np.concatenate([[0.0], rng.uniform(0.5, 1.5, d["z"].size - 1)]))
    t = H.targets_of(d)
    C, P = onp.set_constants(d["x"], d["y"], z, t, d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    fd = onp.spacings(d["x"], d["y"], z)
    dt = onp.cfl_time_step(fd[0], fd[1], fd[2], 0.2, d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    wa = {"f": 3000.0, "source_delay": 1e-6}
    o = onp.OracleSolver(d["x"], d["y"], z, C, P, dt, wave="ricker", wave_args=wa)
    init = {}
    for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old"):
        init[k] = rng.standard_normal(getattr(o, k).shape) * 1e-3
        getattr(o, k)[...] = init[k]
    for k in ("ux", "uy", "uz"):
        getattr(o, k + "_new")[...] = getattr(o, k)
    with H.engine_from_case(d["x"], d["y"], z, d["targets"], d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"],
                            0.2, "ricker", wa, 4, dtype="f64", arith="exact", kernel=kernel) as e:
        assert e.dt == dt
        e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
        e.set_fields(init["ux_old"], init["uy_old"], init["uz_old"], which=1)
        for n in range(4):
            o.step()
            e.run(1)
            for got, key in zip(e.get_fields(), ("ux", "uy", "uz")):
                assert np.array_equal(got, getattr(o, key)), (n, key)
        for got, key in zip(e.get_stress(which=1), ("T1", "T2", "T3", "T4", "T5", "T6")):
            assert np.array_equal(got, getattr(o, key)), key


@pytest.mark.parametrize("kernel", KERNELS)
@pytest.mark.parametrize("name", H.golden_names())
def test_compensated_state_within_tolerance(name, kernel):
    """PHB_COMP (state u, delta): same step algebraically -> fp64 <= 1e-12, fp32 <= 1e-5 of the reference;
    `old` fields read back as u - delta."""
    d = H.load_golden(name)
    for dtype, tol in (("f64", TOL_F64_FAST), ("f32", TOL_F32)):
        with H.engine_from_golden(d, dtype=dtype, arith="comp", kernel=kernel) as e:
            e.run(d["steps"])
            got, old = e.get_fields(), e.get_fields(which=1)
        assert H.rel_l2(got, [d["ux"], d["uy"], d["uz"]]) <= tol, (name, dtype)
        assert H.rel_l2(old, [d["ux_old"], d["uy_old"], d["uz_old"]]) <= 10 * tol, (name, dtype)
