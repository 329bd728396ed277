"""Randomised shapes: 16 seeded small grids with every dimension drawn independently (4 .. 40 points in x and y, 4 .. 150
in z, so one to three z-tiles in fp64, fused and unfused z face, one-launch and ordered faces, both kernels), meshes
non-uniform in all three axes, one inclusion, random initial fields in every entry, a delayed Ricker source --
fp64 EXACT against the C oracle (absorbing y faces) and the NumPy oracle (periodic y), bit for bit."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_edges import _case, _oracle

pytestmark = pytest.mark.gpu


def _shape(seed):
    rng = np.random.default_rng(1000 + seed)
    return (int(rng.integers(4, 41)), int(rng.integers(4, 41)), int(rng.integers(4, 151))), rng


def _init(rng, shape):
    nx, ny, nz = shape
    return {k: rng.standard_normal(s) * 1e-3 for k, s in
            (("ux", (nx - 1, ny, nz)), ("uy", (nx, ny - 1, nz)), ("uz", (nx, ny, nz - 1)),
             ("ux_old", (nx - 1, ny, nz)), ("uy_old", (nx, ny - 1, nz)), ("uz_old", (nx, ny, nz - 1)))}


@pytest.mark.parametrize("seed", range(16))
def test_random_shape_absorbing_bitwise(seed):
    shape, rng = _shape(seed)
    nx, ny, nz = shape
    tg = [(nx * 0.5, ny * 0.4, nz * 0.7, max(1.0, min(nx, ny) * 0.3))]
    case = _case(shape, rng, tg)
    init = _init(rng, shape)
    steps = 6
    ref = None
    for kernel in ("march", "naive"):
        with case.make_engine(steps=steps, dtype="f64", arith="exact", kernel=kernel) as e:
            ids = e.get_material_ids()
            e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
            e.set_fields(init["ux_old"], init["uy_old"], init["uz_old"], which=1)
            e.run(steps)
            got = e.get_fields() + e.get_fields(which=1)
        if ref is None:
            ref = _oracle(case, ids, steps, init)
        for a, k in zip(got, ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old")):
            assert np.array_equal(a, ref[k]), (shape, kernel, k, float(np.abs(a - ref[k]).max()))


@pytest.mark.parametrize("seed", range(8))
def test_random_shape_periodic_bitwise(seed):
    from oracle import fdtd_numpy as onp
    shape, rng = _shape(100 + seed)
    nx, ny, nz = shape
    tg = [(nx * 0.5, ny * 0.1, nz * 0.7, max(1.0, min(nx, ny) * 0.3))]
    case = _case(shape, rng, tg)
    init = _init(rng, shape)
    steps = 6
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
