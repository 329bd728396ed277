"""Every configuration of the marching kernel gives the reference's bits: one / two rows per warp
(PHB_MARCH_RW), the z = -1 absorbing face inside the stencil kernel or as its own kernel (PHB_ZFUSE), the
step split into specialised launches per z-tile class or not (PHB_ZSPLIT),
several z-tiles, y-tiles and x-chunks, grids whose nz makes the fused face possible (nz a multiple of
the vector width) and grids where the host must fall back to the separate face kernel.  fp64 EXACT
arithmetic is compared bit for bit with the C oracle (which follows base_solver.py:245-260, 323-571);
fp32 within 1e-5 (BASELINE.json)."""
import numpy as np
import pytest

from tests import helpers as H
from tests.test_gpu_edges import _case, _oracle

pytestmark = pytest.mark.gpu

VARIANTS = [
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "1"},
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "0"},
    {"PHB_MARCH_RW": "1"},
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "1", "PHB_MARCH_CHUNKS": "3"},
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "1", "PHB_ZSPLIT": "0"},      # fused face, one launch for all z-tiles
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "1", "PHB_ZSPLIT": "1"},      # split step, three parts: k = 0 tile | tiles between | face tile (fp64 default)
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "1", "PHB_ZSPLIT": "2"},      # split step, two parts: face tile | the rest
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "1", "PHB_GRAPH": "0"},
    {"PHB_MARCH_RW": "2", "PHB_ZFUSE": "1", "PHB_FACES_FUSED": "0"},   # ordered x, y, z face launches instead of the one-launch faces kernel
]
SHAPES = [
    ((24, 33, 128), 9),      # fp64: 2 z-tiles, face in the last lane of the second; 3 y-tiles
    ((19, 20, 72), 8),       # fp64: face lane 3 of the second z-tile; fp32: first tile, lane 17
    ((40, 17, 64), 11),      # exactly one fp64 z-tile
    ((12, 30, 66), 7),       # fp64: nz - 1 = 65 sits in lane 0 of the second tile -> host keeps the face kernel
    ((16, 16, 35), 7),       # nz odd -> no fusion in either precision
    ((14, 20, 200), 12),     # fp64: 4 z-tiles (split step: 3 + 1), fp32: 2; enough steps for the CUDA-graph replay of the split
]


@pytest.mark.parametrize("env", VARIANTS, ids=lambda e: ",".join("%s=%s" % (k[4:], v) for k, v in e.items()))
@pytest.mark.parametrize("shape,steps", SHAPES)
def test_march_variants_bitwise_vs_c_oracle(shape, steps, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    rng = np.random.default_rng(sum(shape) + 1)
    nx, ny, nz = shape
    tg = [(nx * 0.45, ny * 0.55, nz * 0.7, min(nx, ny) * 0.25)]
    case = _case(shape, rng, tg)
    init = {k: rng.standard_normal(s) * 1e-3 for k, s in
            (("ux", (nx - 1, ny, nz)), ("uy", (nx, ny - 1, nz)), ("uz", (nx, ny, nz - 1)),
             ("ux_old", (nx - 1, ny, nz)), ("uy_old", (nx, ny - 1, nz)), ("uz_old", (nx, ny, nz - 1)))}
    with case.make_engine(steps=steps, dtype="f64", arith="exact", kernel="march") as e:
        assert e.info()["kernel"] == "march_tma"
        ids = e.get_material_ids()
        e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
        e.set_fields(init["ux_old"], init["uy_old"], init["uz_old"], which=1)
        e.run(steps)
        got = e.get_fields() + e.get_fields(which=1)
    ref = _oracle(case, ids, steps, init)
    for a, k in zip(got, ("ux", "uy", "uz", "ux_old", "uy_old", "uz_old")):
        assert np.array_equal(a, ref[k]), (shape, env, k, float(np.abs(a - ref[k]).max()))
    for dtype, arith, tol in (("f64", "fast", 1e-12), ("f32", "fast", 1e-5)):
        with case.make_engine(steps=steps, dtype=dtype, arith=arith, kernel="march") as e:
            e.set_fields(init["ux"], init["uy"], init["uz"], which=0)
            e.set_fields(init["ux_old"], init["uy_old"], init["uz_old"], which=1)
            e.run(steps)
            assert H.rel_l2(e.get_fields(), [ref["ux"], ref["uy"], ref["uz"]]) <= tol, (shape, env, dtype)
