"""Multi-rank parity: sharded scans, partial-table merge, partition + row exchange, joins -- on two GPUs where the box has
them, otherwise as two ranks SHARING the one GPU (the rows still cross process boundaries through the library's
peer-memory exchange; see tests/multi_gpu_worker.py)."""

from __future__ import annotations

import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def _gpu_count() -> int:
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize(("world", "peer"), [(2, "1"), (2, "0")], ids=["peer-memory-exchange", "nccl-exchange"])
def test_sharded_queries_match_oracle(tmp_path, world, peer):
    if _gpu_count() < 1:
        pytest.skip("needs a GPU")
    if peer == "0" and _gpu_count() < world:
        pytest.skip("the NCCL exchange needs one GPU per rank")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    worker = Path(__file__).parent / "multi_gpu_worker.py"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(worker), str(tmp_path)]
    import os

    env = dict(os.environ, MINISPARK_PEER_EXCHANGE=peer)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0, out.stdout[-4000:] + out.stderr[-4000:]
    assert f"multi-gpu ok {world}" in out.stdout
