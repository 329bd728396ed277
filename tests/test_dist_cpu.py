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


def test_connect_runs_the_same_collectives_on_every_rank_when_one_fails():
    """Engine.connect (the halo set-up): a rank whose CUDA-IPC export fails must still take part in both all-gathers, so
    that every rank falls back to NCCL together instead of pairing mismatched collectives.  Two "ranks" =
    two threads with a barrier-backed all-gather and a stub in place of the ctypes calls."""
    import threading
    from phonomena_b200 import _lib

    world = 2
    barrier = threading.Barrier(world)
    slots, lock, log = {}, threading.Lock(), {0: [], 1: []}

    def make_allgather(rank):
        n = [0]

        def allgather(obj):
            k = n[0]
            n[0] += 1
            with lock:
                slots.setdefault(k, {})[rank] = obj
            barrier.wait(timeout=10)
            out = [slots[k][r] for r in range(world)]
            log[rank].append(("allgather", k))
            return out
        return allgather

    def broadcast_from(rank):
        ag = make_allgather(100 + rank)      # not used: broadcast is modelled with the same barrier
        def broadcast(obj):
            with lock:
                if rank == 0:
                    slots["uid"] = obj
            barrier.wait(timeout=10)
            return slots["uid"]
        return broadcast

    class Stub(_lib.Engine):
        def __init__(self, rank, fail_export):
            self.rank, self.fail_export, self.calls = rank, fail_export, []
        def p2p_export(self):
            self.calls.append("export")
            if self.fail_export:
                raise _lib.PhbError("cudaIpcGetMemHandle failed")
            return (b"h" * 256, 24)
        def p2p_import(self, rank, nranks, exports):
            self.calls.append("import")
        def comm_init(self, uid, rank, nranks):
            self.calls.append(("nccl", uid))
        def close(self):
            pass

    import phonomena_b200._lib as L
    orig = L.comm_unique_id
    L.comm_unique_id = lambda: b"u" * 128
    try:
        results, engines = {}, {r: Stub(r, fail_export=(r == 1)) for r in range(world)}

        def run(rank):
            results[rank] = engines[rank].connect(rank, world, make_allgather(rank), broadcast_from(rank), mode="p2p")
        ts = [threading.Thread(target=run, args=(r,)) for r in range(world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join(20)
        assert not any(t.is_alive() for t in ts), "connect deadlocked"
    finally:
        L.comm_unique_id = orig
    assert results == {0: "nccl", 1: "nccl"}
    assert log[0] == log[1] == [("allgather", 0), ("allgather", 1)]          # the same two collectives on both ranks
    assert engines[0].calls[-1] == ("nccl", b"u" * 128) and engines[1].calls[-1] == ("nccl", b"u" * 128)
    assert "import" not in engines[0].calls and "import" not in engines[1].calls      # nobody imports a missing export
