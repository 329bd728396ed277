"""Run-time self-checks of the multi-GPU path (no oracle involved: CUDA slabs against the CUDA
single-GPU run of the same grid on the same device -- same kernel, same arithmetic order, so the
results must be BIT-IDENTICAL; SURVEY 8d "Multi-GPU: P-GPU result bit-identical to 1-GPU").

Used by bench.py before the timed region of every N > 1 run (the `parity` object of its JSON line)
and by tools/multi_check.py / tests/test_gpu_multi.py.
"""
from __future__ import annotations

import numpy as np

from . import hostmath as hm
from .workloads import crystal_case

DEFAULT_MODES = (("f64", "fast"), ("f64", "exact"), ("f32", "fast"))


def _initial_fields(case, world, seed=7):
    """Random displacements (|u| ~ 1e-3) that vanish within 3 planes of every slab boundary: the ghost planes of
    a freshly created slab context are zero, so the initial state must be; two steps later real data crosses
    every boundary in both directions (the left-wall source alone would reach the far slabs only as underflow)."""
    nx, ny, nz = case.shape
    rng = np.random.default_rng(seed)
    mask = np.zeros(nx)
    for x0, nxl in hm.split_slabs(nx, world):
        mask[x0 + 3:x0 + nxl - 3] = 1.0
    shp = ((nx - 1, ny, nz), (nx, ny - 1, nz), (nx, ny, nz - 1))
    cur = [rng.standard_normal(s) * 1e-3 * mask[:s[0], None, None] for s in shp]
    old = [a * 0.97 for a in cur]
    return cur, old


def slabs_vs_single(rank, world, device, allgather, broadcast, grid=None, steps=40, modes=DEFAULT_MODES, halo=None,
                    kernel="march"):
    """Every rank steps its x-slab of a small phononic crystal through the halo exchange the run will
    use, then the whole grid alone on its own GPU, and compares the owned planes with np.array_equal.
    Returns {"slabs_bit_identical": bool, "halo": mode, "grid": [...], "steps": n, "modes": [...]}
    (the same dict on every rank).  `grid` defaults to (24 * world, 72, 80), at least 96 planes: slabs of
    >= 12 planes; the state starts from random fields so every slab boundary carries data from step 2 on.  Both runs
    use the production (marching) kernel: FAST arithmetic is bit-identical only within one kernel."""
    nx, ny, nz = grid or (max(96, 24 * world), 72, 80)
    case = crystal_case(nx, ny, nz)
    x0, nxl = hm.split_slabs(nx, world)[rank]
    cur, old = _initial_fields(case, world)
    ok, used = True, None
    for dtype, arith in modes:
        e = case.make_engine(steps=steps, x0=x0, nxl=nxl, dtype=dtype, arith=arith, device=device, kernel=kernel)
        used = e.connect(rank, world, allgather, broadcast, mode=halo)
        sl = [slice(x0, x0 + e.planes(c)) for c in range(3)]
        e.set_fields(*[a[s] for a, s in zip(cur, sl)], which=0)
        e.set_fields(*[a[s] for a, s in zip(old, sl)], which=1)
        e.run(steps)
        e.sync()
        mine = e.get_fields()
        e.close()
        f = case.make_engine(steps=steps, dtype=dtype, arith=arith, device=device, kernel=kernel)
        f.set_fields(*cur, which=0)
        f.set_fields(*old, which=1)
        f.run(steps)
        full = f.get_fields()
        f.close()
        same = all(np.array_equal(a, b[x0:x0 + a.shape[0]]) for a, b in zip(mine, full))
        # data really reached the slab boundary: the first and last owned ux planes started as zeros
        moved = bool(np.any(mine[0][0] != 0) and np.any(mine[0][-1] != 0))
        res = allgather((bool(same), bool(moved)))
        ok = ok and all(r[0] and r[1] for r in res)
    return {"slabs_bit_identical": bool(ok), "halo": used, "grid": [nx, ny, nz], "steps": int(steps),
            "modes": ["%s/%s" % m for m in modes]}
