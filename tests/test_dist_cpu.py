"""Host-side logic of the N>1 path on CPU (gloo, world_size 2): slab assignment from the torchrun
environment, the neighbour/plane plan of the halo exchange, id broadcast."""
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    from phonomena_b200 import hostmath as hm
    from phonomena_b200.solver_b200 import slab_from_env
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    nx, ny, nz = 37, 5, 4
    x0, nxl, r, n = slab_from_env(nx)
    assert (r, n) == (rank, world) and (x0, nxl) == hm.split_slabs(nx, world)[rank]
    # global field, each rank holds planes [x0-1, x0+nxl] (ghosts included) like the device layout
    glob = np.arange(nx * ny * nz, dtype=np.float64).reshape(nx, ny, nz)
    loc = np.zeros((nxl + 2, ny, nz)); loc[1:nxl + 1] = glob[x0:x0 + nxl]
    ops = []
    for peer, send_l, recv_l in hm.halo_plan(rank, world, nxl):
        s = torch.from_numpy(np.ascontiguousarray(loc[send_l])); rcv = torch.zeros(ny, nz, dtype=torch.float64)
        ops.append((dist.isend(s, peer), None)); ops.append((dist.irecv(rcv, peer), (recv_l, rcv)))
    for w, tgt in ops:
        w.wait()
        if tgt: loc[tgt[0]] = tgt[1].numpy()
    lo, hi = max(x0 - 1, 0), min(x0 + nxl + 1, nx)
    assert np.array_equal(loc[lo - (x0 - 1):hi - (x0 - 1)], glob[lo:hi]), rank
    uid = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    assert isinstance(uid[0], bytes) and len(uid[0]) == 128
    dist.destroy_process_group()
    print("rank", rank, "ok")
''') % ROOT


def test_slab_plan_two_ranks_gloo(tmp_path):
    p = tmp_path / "w.py"
    p.write_text(SCRIPT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(p)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout
