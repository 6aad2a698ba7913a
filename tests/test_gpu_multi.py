"""Real-NCCL parity of the sharded engines: launches ``tests/dist_gpu_worker.py`` under
``torch.distributed.run`` with one process per GPU.  Needs >= 2 GPUs (``gpurun --gpus 2``); on the
one-GPU box of the round-end run it is skipped (the in-process emulations in
``test_gpu_parity.py`` and the gloo tests in ``test_dist_cpu.py`` cover the same math there).
The log of the last multi-GPU run is kept under ``profiles/``."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_engines_real_nccl(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert f"DIST PARITY OK world={world}" in r.stdout, tail
