"""The REAL multi-process row-sharded path (one process per GPU, NCCL process group, CUDA-IPC peer
mailboxes) against the single-store result: tools/check_sharded.py under torch.distributed.run with
2 ranks.  Needs 2 visible GPUs (skipped on a one-GPU box, where tests/test_gpu_sharded.py emulates
the ranks on one device and tests/test_gpu_mstore.py runs the same kernels behind one handle)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_process_sharded_search_equals_single_store(native):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 visible GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "check_sharded.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-4000:]
    assert "SHARDED_OK" in r.stdout and "MISMATCH" not in r.stdout
