"""Multi-process path on CPU: world_size 2 over gloo (the GPU box runs the same code over
NCCL).  Channels shard with no data-path collective; all-gather only reassembles."""
from __future__ import annotations

import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT
from torchfx_b200.dist import shard_bounds


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_partition():
    for C in (1, 7, 8, 1024, 1000):
        for P in (1, 2, 3, 8):
            spans = [shard_bounds(C, P, r) for r in range(P)]
            assert spans[0][0] == 0 and spans[-1][1] == C
            assert all(spans[i][1] == spans[i + 1][0] for i in range(P - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


@pytest.mark.timeout(300)
def test_world_size_2_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=280, env=env)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert "DIST_OK" in p.stdout
