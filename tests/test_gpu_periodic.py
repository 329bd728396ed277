"""SURVEY 8f row 4 -- periodic y boundaries (zero Bloch phase): the CUDA path with cfg bc_y = "periodic" against
fixtures made by the unmodified reference solver with its archived apply_T_pbc / apply_u_pbc stubs switched on
(tests/golden/periodic_y_*.npz, oracle/gen_golden.periodic_y_solver), and against the NumPy oracle's periodic mode
(pinned to the same fixtures on the CPU) on grids with several y-tiles, z-tiles and random initial fields."""
import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kernel", ["naive", "march"])
@pytest.mark.parametrize("name", H.periodic_names())
def test_periodic_y_bitwise_vs_reference_stubs(name, kernel):
    d = H.load_golden(name)
    snaps = sorted(int(s) for s in d["snap_steps"])
    with H.engine_from_golden(d, dtype="f64", arith="exact", kernel=kernel, bc_y="periodic") as e:
        assert np.array_equal(e.get_material_ids(), d["ids"])
        done = 0
        for n in snaps:
            e.run(n - done)
            done = n
            ux, uy, uz = e.get_fields()
            assert np.array_equal(ux[:, :, 0], d["snap_ux_%d" % n]) and np.array_equal(uy[:, :, 0], d["snap_uy_%d" % n]), (name, n)
            assert np.array_equal(uz[:, :, 0], d["snap_uz_%d" % n]), (name, n)
        e.run(d["steps"] - done)
        for got, key in zip(e.get_fields(), ("ux", "uy", "uz")):
            assert np.array_equal(got, d[key]), (name, key, float(np.abs(got - d[key]).max()))
        for got, key in zip(e.get_fields(which=1), ("ux_old", "uy_old", "uz_old")):
            assert np.array_equal(got, d[key]), (name, key)
        for got, key in zip(e.get_stress(which=1), ("T1", "T2", "T3", "T4", "T5", "T6")):
            assert np.array_equal(got, d[key]), (name, key)
    ref = [d["ux"], d["uy"], d["uz"]]
    for dtype, arith, tol in (("f64", "fast", 1e-12), ("f32", "fast", 1e-5)):
        with H.engine_from_golden(d, dtype=dtype, arith=arith, kernel=kernel, bc_y="periodic") as e:
            e.run(d["steps"])
            assert H.rel_l2(e.get_fields(), ref) <= tol, (name, dtype)


@pytest.mark.parametrize("shape,steps", [((26, 45, 70), 10), ((18, 16, 200), 8), ((9, 7, 5), 9)])
def test_periodic_y_random_fields_vs_oracle(shape, steps):
    """Several y-tiles / z-tiles (fused z face, split step), non-uniform mesh in all axes, an inclusion cutting the
    periodic rows, random initial fields in every entry (so the rows nothing writes in this mode must keep u_new == u)."""
    from oracle import fdtd_numpy as onp
    from tests.test_gpu_edges import _case
    rng = np.random.default_rng(sum(shape) + 7)
    nx, ny, nz = shape
    tg = [(nx * 0.5, ny * 0.15, nz * 0.8, min(nx, ny) * 0.3)] if min(nx, ny) >= 9 else []
    case = _case(shape, rng, tg)
    init = {k: rng.standard_normal(s) * 1e-3 for k, s in
            (("ux", (nx - 1, ny, nz)), ("uy", (nx, ny - 1, nz)), ("uz", (nx, ny, nz - 1)),
             ("ux_old", (nx - 1, ny, nz)), ("uy_old", (nx, ny - 1, nz)), ("uz_old", (nx, ny, nz - 1)))}
    C, P = onp.set_constants(case.x, case.y, case.z, onp.make_targets(case.targets.tolist()), case.prim_c, case.prim_p, case.sec_c, case.sec_p)
    o = onp.OracleSolver(case.x, case.y, case.z, C, P, case.dt, wave=case.wave, wave_args=case.wave_args, bc_y="periodic")
    for k, a in init.items():
        getattr(o, k)[...] = a
    for k in ("ux", "uy", "uz"):
        getattr(o, k + "_new")[...] = getattr(o, k)
    o.run(steps)
    for kernel in ("march", "naive"):
        with case.make_engine(steps=steps, dtype="f64", arith="exact", kernel=kernel, bc_y="periodic") as e:
            e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
            e.set_fields(init["ux_old"], init["uy_old"], init["uz_old"], which=1)
            e.run(steps)
            got = e.get_fields() + e.get_fields(which=1)
        for a, k in zip(got, ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old")):
            assert np.array_equal(a, getattr(o, k)), (shape, kernel, k, float(np.abs(a - getattr(o, k)).max()))


def test_periodic_y_plugin_cfg_and_errors(tmp_path):
    from phonomena_b200 import _lib
    from tests.test_gpu_plugin import fake_from_golden, make_solver
    d = H.load_golden("periodic_y_homog_24x12x10")
    s = make_solver(d, tmp_path, write_mode="off", bc_y="periodic")
    s.init(*fake_from_golden(d), d["steps"])
    s.run()
    for got, key in zip(s.fields(), ("ux", "uy", "uz")):
        assert np.array_equal(got, d[key]), key
    with pytest.raises(_lib.PhbError, match="compensated"):
        H.engine_from_golden(d, arith="compensated", dtype="f32", bc_y="periodic")
    with pytest.raises(KeyError):
        H.engine_from_golden(d, bc_y="bloch")


# ---- Bloch phase (beyond the reference: its archived stubs are the phase-0 case) -------------------------------------
def _bloch_engines(d, phase, dtype="f64", arith="exact", kernel="auto"):
    re = H.engine_from_golden(d, dtype=dtype, arith=arith, kernel=kernel, bc_y="periodic")
    im = H.engine_from_golden(d, dtype=dtype, arith=arith, kernel=kernel, bc_y="periodic")
    im.set_source_table(None)
    re.bloch_pair(im, phase)
    return re, im


@pytest.mark.parametrize("kernel", ["naive", "march"])
def test_bloch_phase_zero_is_the_reference_periodic_case(kernel):
    """phase = 0: the real part of a Bloch pair is bit-identical to the reference-with-stubs fixture, the imaginary part
    stays exactly zero."""
    d = H.load_golden("periodic_y_crystal_40x18x14")
    re, im = _bloch_engines(d, 0.0, kernel=kernel)
    re.run(d["steps"])
    re.sync()
    for got, key in zip(re.get_fields(), ("ux", "uy", "uz")):
        assert np.array_equal(got, d[key]), key
    assert all(not a.any() for a in im.get_fields())
    with pytest.raises(Exception, match="imaginary part"):
        im.run(1)
    im.close()
    re.close()


@pytest.mark.parametrize("phase", [0.7, np.pi, -2.1])
def test_bloch_phase_vs_oracle(phase):
    """phase != 0: the device against oracle.fdtd_numpy.BlochOracle (which DEFINES these semantics; parity unpinned by
    the reference, which has no phase) -- EXACT arithmetic bit for bit in both parts, FAST <= 1e-12, fp32 <= 1e-5; and the
    Bloch relation itself: the periodic rows of the complex field differ by exp(-+ i phase)."""
    from oracle import fdtd_numpy as onp
    d = H.load_golden("periodic_y_crystal_40x18x14")
    steps = 90
    t = H.targets_of(d)
    C, P = onp.set_constants(d["x"], d["y"], d["z"], t, d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    o = onp.BlochOracle(d["x"], d["y"], d["z"], C, P, d["dt"], phase, wave=d["wave"], wave_args=d["wave_args"]).run(steps)
    ref_re, ref_im = [o.re.ux, o.re.uy, o.re.uz], [o.im.ux, o.im.uy, o.im.uz]
    assert sum(float(np.abs(a).sum()) for a in ref_im) > 0           # the phase really couples the parts
    for kernel in ("march", "naive"):
        re, im = _bloch_engines(d, phase, kernel=kernel)
        re.run(steps)
        re.sync()
        got_re, got_im = re.get_fields(), im.get_fields()
        im.close(); re.close()
        for a, b, n in zip(got_re + got_im, ref_re + ref_im, ("ux", "uy", "uz", "ux_im", "uy_im", "uz_im")):
            assert np.array_equal(a, b), (kernel, n, float(np.abs(a - b).max()))
    z = got_re[0] + 1j * got_im[0]                                    # ux: row 0 = exp(-i phase) * row ny-2 (off the z face)
    assert np.allclose(z[:, 0, :-1], np.exp(-1j * phase) * z[:, -2, :-1], rtol=1e-12, atol=1e-300)
    for dtype, arith, tol in (("f64", "fast", 1e-12), ("f32", "fast", 1e-5)):
        re, im = _bloch_engines(d, phase, dtype=dtype, arith=arith)
        re.run(steps)
        re.sync()
        assert H.rel_l2(re.get_fields() + im.get_fields(), ref_re + ref_im) <= tol, (dtype, arith)
        im.close(); re.close()


def test_bloch_plugin_cfg(tmp_path):
    from oracle import fdtd_numpy as onp
    from tests.test_gpu_plugin import fake_from_golden, make_solver
    d = H.load_golden("periodic_y_homog_24x12x10")
    s = make_solver(d, tmp_path, write_mode="off", bc_y="bloch", bloch_phase=1.1)
    s.init(*fake_from_golden(d), 40)
    s.run()
    C, P = onp.set_constants(d["x"], d["y"], d["z"], H.targets_of(d), d["prim_c"], d["prim_p"], d["sec_c"], d["sec_p"])
    o = onp.BlochOracle(d["x"], d["y"], d["z"], C, P, d["dt"], 1.1, wave=d["wave"], wave_args=d["wave_args"]).run(40)
    for a, b in zip(s.fields() + s.fields_imag(), [o.re.ux, o.re.uy, o.re.uz, o.im.ux, o.im.uy, o.im.uz]):
        assert np.array_equal(a, b)
