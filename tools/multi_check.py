"""Multi-GPU parity check, launched by torchrun (one process per GPU):
every rank steps its x-slab with NCCL halo exchange; the slabs must be BIT-IDENTICAL to the
same planes of a single-GPU run (same kernel, same arithmetic order)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from phonomena_b200 import _lib, hostmath as hm
from phonomena_b200.workloads import crystal_case


def spectrum_check(rank, world, local, nx, ny, nz, steps):
    """Plugin level: Solver.spectrum over slabs (partial 2-D transforms summed over the ranks, the 1-D one
    taken from the owning rank) against the same call on a single-GPU run of the whole grid."""
    from phonomena_b200.solver_b200 import Solver
    g, m = crystal_case(nx, ny, nz).as_grid_material()
    out, files = [], []
    for slabs in (True, False):
        s = Solver()
        s.cfg.update({"wave": "sin", "wave_args": {"f": 100}, "write_mode": "thread", "record": "surface", "arith": "exact",
                      "device": local, "slabs_from_env": slabs, "record_every": 4,
                      "probes": [{"u": "ux", "y": ny // 3, "z": 0}, {"u": "uz", "y": ny // 2, "z": 2}]})
        s.init(g, m, steps)
        s.run()
        out.append([s.spectrum("ux", 0, ny // 3), s.spectrum("uz", 2, ny // 2), s.spectrum("ux", 0, ny // 3, x_index=1),
                    s.spectrum("uz", 2, ny // 2, x_index=-world)])
        s._close_engine()
        files.append(s.file)
    same_file = 1.0
    if rank == 0:      # the merged slab file against the single-GPU file: same schema, same frames
        from phonomena_b200.h5lite import H5Reader
        a, b = H5Reader(files[0]), H5Reader(files[1])
        same_file = float(all(a.shape(k) == b.shape(k) and np.array_equal(a.read(k), b.read(k)) for k in ("ux", "uy", "uz", "density"))
                          and a.attrs["frames_written"] == b.attrs["frames_written"] == steps // 4 and a.attrs["nxl"] == nx
                          and not os.path.exists(files[0] + ".rank0"))
    for f in files:
        if os.path.exists(f):
            os.remove(f)
    err = 0.0
    for a, b in zip(*out):
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and a[2].shape == b[2].shape
        assert np.linalg.norm(b[2]) > 0          # probes sit where the wave already is
        err = max(err, float(np.linalg.norm(a[2] - b[2]) / np.linalg.norm(b[2])))
    t = torch.tensor([err, 1.0 - same_file], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"check": "plugin over slabs vs single GPU", "world": world, "spectrum_max_rel_l2": float(t[0]),
                          "merged_file_identical": int(t[1] == 0)}), flush=True)
    return float(t[0]) <= 1e-12 and float(t[1]) == 0


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nx, ny, nz = [int(v) for v in os.environ.get("PHB_MC_GRID", "96,72,80").split(",")]
    steps = int(os.environ.get("PHB_MC_STEPS", "40"))
    def allgather(obj):
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out

    def broadcast(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    halo = os.environ.get("PHB_HALO", "p2p")
    from phonomena_b200 import selfcheck
    ok = True
    # (1) what bench.py runs before every N > 1 timed region: random initial fields, the production halo path
    for hm_ in dict.fromkeys([halo, "fused", "nccl"]):      # the production mode first, then the in-kernel push and the NCCL fallback
        res = selfcheck.slabs_vs_single(rank, world, local, allgather, broadcast, halo=hm_)
        if rank == 0:
            print(json.dumps({"check": "selfcheck.slabs_vs_single", "world": world, **res}), flush=True)
        ok = ok and res["slabs_bit_identical"] and res["halo"] == hm_
    # (2) source-driven runs from zero fields, marching kernel over the chosen halo mode and the per-cell kernel over NCCL
    for dtype, arith, kernel in (("f64", "fast", "march"), ("f64", "exact", "march"), ("f32", "fast", "march"), ("f64", "fast", "naive")):
        case = crystal_case(nx, ny, nz)
        x0, nxl = hm.split_slabs(nx, world)[rank]
        e = case.make_engine(steps=steps, x0=x0, nxl=nxl, dtype=dtype, arith=arith, device=local, kernel=kernel)
        mode = e.connect(rank, world, allgather, broadcast, mode=halo if kernel == "march" else "nccl")
        e.run(steps)
        e.sync()
        mine = e.get_fields()
        e.close()
        # single-GPU run of the whole grid on every rank's own device, compared on the owned planes
        f = case.make_engine(steps=steps, dtype=dtype, arith=arith, device=local, kernel=kernel)
        f.run(steps)
        full = f.get_fields()
        f.close()
        same = all(np.array_equal(a, b[x0:x0 + a.shape[0]]) for a, b in zip(mine, full))
        nrm = float(sum(np.sum(a.astype(np.float64) ** 2) for a in mine))
        t = torch.tensor([1.0 if same else 0.0, nrm], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        if rank == 0:
            print(json.dumps({"halo": mode, "dtype": dtype, "arith": arith, "kernel": kernel, "ranks_identical": int(t[0]), "world": world,
                              "energy": float(t[1])}), flush=True)
        ok = ok and int(t[0]) == world and float(t[1]) > 0
    ok = spectrum_check(rank, world, local, nx, ny, nz, int(os.environ.get("PHB_MC_LONG", "1000"))) and ok      # the wave crosses every slab
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
