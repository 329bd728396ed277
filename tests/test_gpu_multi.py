"""x-slab decomposition with NCCL halo exchange: P-GPU result bit-identical to 1-GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    from phonomena_b200 import _lib
    try:
        return _lib.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_bit_identical_to_single_gpu(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tools", "multi_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('"ranks_identical": %d' % world) == 4, r.stdout
