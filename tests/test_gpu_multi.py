"""GPU suite (-m gpu), multi-GPU part: needs >= 2 GPUs on the box (skipped otherwise; run with `gpurun --gpus 2|8`).
Spawns tests/multi_gpu_worker.py under torch.distributed.run: row-sharded GEMV whose all-gather is fused into the kernel
epilogue (peer stores over NVLink, no collective launch); gathered == un-sharded bit for bit on every rank, incl. the
22/21-tile remainder of 11008 rows at world 8; plus a decode sequence (resident chain kernel) whose ops store their rows to the
peers from the kernel's epilogue (tmac_b200_seq_peer_outputs)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _worlds():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return sorted({w for w in (2, 4, 8) if w <= n})


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_gemv_with_fused_gather_equals_unsharded(world):
    if world not in _worlds():
        pytest.skip("needs %d GPUs" % world)
    port = 29500 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "multi_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK ") == 5, r.stdout[-2000:]
