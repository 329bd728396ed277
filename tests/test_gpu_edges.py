"""Edge cases of the hot path on the GPU: ragged grids (no dimension a multiple of any tile size,
several z-tiles, several x-chunks, partial last y-tile), the minimum grid, runs without source or
inclusions, non-uniform meshes in all three axes -- every array compared bit for bit with the C
oracle in EXACT arithmetic -- and the error behaviour of the C ABI."""
import os

import numpy as np
import pytest

from tests import helpers as H

pytestmark = pytest.mark.gpu

PROPS = None


def _case(shape, rng=None, targets=(), wave="ricker", wave_args=None, courant=0.3):
    from phonomena_b200.workloads import Case
    nx, ny, nz = shape
    if rng is None:
        x, y, z = (np.arange(n, dtype=np.float64) for n in shape)
    else:
        x, y, z = (np.cumsum(np.concatenate([[0.0], rng.uniform(0.4, 1.6, n - 1)])) for n in shape)
    return Case(x, y, z, np.asarray(targets, np.float32).reshape(-1, 4), courant=courant, wave=wave,
                wave_args=wave_args or {"f": 4000.0, "source_delay": 5e-5})


def _oracle(case, ids, steps, init=None):
    from oracle import fdtd_c
    o = fdtd_c.COracle(case.x, case.y, case.z, ids, [case.prim_c, case.sec_c], [case.prim_p, case.sec_p], case.dt,
                       wave=case.wave or "sin", wave_args=case.wave_args)
    if init is not None:
        for k, a in init.items():
            getattr(o, k)[...] = a
        for k in ("ux", "uy", "uz"):
            getattr(o, k + "_new")[...] = getattr(o, k)
    o.run(steps)
    out = {k: getattr(o, k).copy() for k in ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old")}
    o.close()
    return out


RAGGED = [
    ((5, 4, 4), 6, {}),                                   # minimum grid
    ((9, 7, 5), 9, {}),
    ((37, 29, 67), 12, {}),                               # fp64: 2 z-tiles (nzp = 96)
    ((21, 45, 131), 10, {"PHB_MARCH_CHUNKS": "3"}),       # 3 z-tiles in fp64, 4 y-tiles, 3 x-chunks
    ((70, 15, 33), 14, {"PHB_MARCH_CHUNKS": "5"}),        # nz just over one 32-pitch; chunks of 14 planes
    ((18, 16, 200), 8, {"PHB_MARCH_R": "8"}),             # the 8-row block variant
]


@pytest.mark.parametrize("shape,steps,env", RAGGED)
def test_ragged_grids_bitwise_vs_c_oracle(shape, steps, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(sum(shape))
    nx, ny, nz = shape
    tg = [(nx * 0.5, ny * 0.5, nz * 0.6, min(nx, ny) * 0.22)] if min(nx, ny) >= 9 else []
    case = _case(shape, rng, tg)
    init = {k: rng.standard_normal(s) * 1e-3 for k, s in
            (("ux", (nx - 1, ny, nz)), ("uy", (nx, ny - 1, nz)), ("uz", (nx, ny, nz - 1)),
             ("ux_old", (nx - 1, ny, nz)), ("uy_old", (nx, ny - 1, nz)), ("uz_old", (nx, ny, nz - 1)))}
    for kernel in ("march", "naive"):
        with case.make_engine(steps=steps, dtype="f64", arith="exact", kernel=kernel) as e:
            ids = e.get_material_ids()
            e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
            e.set_fields(init["ux_old"], init["uy_old"], init["uz_old"], which=1)
            e.run(steps)
            got = e.get_fields() + e.get_fields(which=1)
        ref = _oracle(case, ids, steps, init)
        for a, k in zip(got, ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old")):
            assert np.array_equal(a, ref[k]), (shape, kernel, k, float(np.abs(a - ref[k]).max()))
    with case.make_engine(steps=steps, dtype="f32", arith="fast", kernel="march") as e:
        e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
        e.set_fields(init["ux_old"], init["uy_old"], init["uz_old"], which=1)
        e.run(steps)
        assert H.rel_l2(e.get_fields(), [ref["ux"], ref["uy"], ref["uz"]]) <= 1e-5


def test_no_source_no_inclusions_stays_zero_and_free_run():
    case = _case((24, 20, 40))
    case.wave = None
    with case.make_engine(steps=5, dtype="f64", arith="exact") as e:
        e.run(5)
        assert not any(np.any(a) for a in e.get_fields())          # nothing in, nothing out
        assert not np.any(e.get_material_ids())
    # free evolution from random fields without a source: the pre-source line logic must not interfere
    rng = np.random.default_rng(5)
    nx, ny, nz = case.shape
    init = {k: rng.standard_normal(s) * 1e-3 for k, s in
            (("ux", (nx - 1, ny, nz)), ("uy", (nx, ny - 1, nz)), ("uz", (nx, ny, nz - 1)))}
    outs = []
    for kernel in ("march", "naive"):
        with case.make_engine(steps=7, dtype="f64", arith="exact", kernel=kernel) as e:
            e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
            e.set_fields(init["ux"], init["uy"], init["uz"], which=1)
            e.run(7)
            outs.append(e.get_fields())
    assert all(np.array_equal(a, b) for a, b in zip(*outs)) and np.any(outs[0][2][0])


def test_c_abi_error_behaviour():
    from phonomena_b200 import _lib
    with pytest.raises(_lib.PhbError, match="at least 4 points"):
        _lib.Engine(3, 8, 8, 1e-5)
    with pytest.raises(_lib.PhbError, match="bad slab"):
        _lib.Engine(8, 8, 8, 1e-5, x0=4, nxl=8)
    with _lib.Engine(8, 8, 8, 1e-5) as e:
        with pytest.raises(_lib.PhbError, match="phb_set_spacing not called"):
            e.run(1)
        one = np.ones(7)
        e.set_spacing(one, one, one, one[:6], one[:6], one[:6])
        with pytest.raises(_lib.PhbError, match="material not set"):
            e.run(1)
        with pytest.raises(_lib.PhbError, match="non-positive"):
            e.set_spacing(one * 0, one, one, one[:6], one[:6], one[:6])
        with pytest.raises(_lib.PhbError, match="nmat must be"):
            e.set_material_table([np.eye(6)] * 16, [1.0] * 16)
        with pytest.raises(_lib.PhbError, match="not positive"):
            e.set_material_table([np.eye(6)], [0.0])
        e.set_material_table([np.eye(6), np.eye(6)], [1.0, 2.0])
        with pytest.raises(_lib.PhbError, match="expected"):
            e.set_material_ids(np.zeros((3, 8, 8), np.uint8))
        with pytest.raises(_lib.PhbError, match="material id 3"):      # id 3 is not in the 2-entry table
            e.set_material_ids(np.full((8, 8, 8), 3, np.uint8))
        e.set_material_ids(np.zeros((8, 8, 8), np.uint8))
        with pytest.raises(_lib.PhbError, match="phb_set_abc not called"):
            e.run(1)
        e.set_abc([0.0] * 8)
        e.set_source_table(np.zeros(2))
        e.run(2)
        with pytest.raises(_lib.PhbError, match="source table covers"):
            e.run(1)
        with pytest.raises(_lib.PhbError, match="shape"):
            e.set_fields(np.zeros((8, 8, 8)), None, None)
        assert e.steps_done == 2 and e.launch_count > 0


def test_random_two_material_medium_every_stencil_class():
    """Per-cell random material (not cylinders): exercises (nearly) all 2^7 stencil classes plus their
    boundary variants; uploaded with phb_set_material_ids.  EXACT arithmetic vs the C oracle, bit for bit,
    with kernel = auto (marching kernel, or the naive one if the class table outgrows shared memory)."""
    from phonomena_b200 import _lib, hostmath as hm
    from oracle import fdtd_c
    rng = np.random.default_rng(11)
    case = _case((20, 18, 70), rng)
    nx, ny, nz = case.shape
    ids = (rng.random((nx, ny, nz)) < 0.4).astype(np.uint8)
    ids[:3, :3, :3] = 0            # corner cell primary (Mur coefficients)
    steps = 9
    o = fdtd_c.COracle(case.x, case.y, case.z, ids, [case.prim_c, case.sec_c], [case.prim_p, case.sec_p], case.dt,
                       wave="ricker", wave_args=case.wave_args)
    o.run(steps)
    for kernel in ("auto", "naive"):
        e = _lib.Engine(nx, ny, nz, case.dt, dtype="f64", arith="exact", kernel=kernel)
        e.set_spacing(*case.sp)
        e.set_material_table([case.prim_c, case.sec_c], [case.prim_p, case.sec_p])
        e.set_material_ids(ids)
        assert np.array_equal(e.get_material_ids(), ids)
        e.set_abc(hm.abc_coefficients(case.prim_c, case.prim_p, case.dt, *case.sp))
        e.set_source_table(hm.source_table("ricker", steps, case.dt, case.wave_args))
        e.run(steps)
        for a, k in zip(e.get_fields(), ("ux", "uy", "uz")):
            assert np.array_equal(a, getattr(o, k)), (kernel, k)
        e.close()
    o.close()
