"""Multi-GPU schedule on real GPUs over NCCL (needs >= 2 devices; skipped otherwise): both exchange paths (fused GEMM ->
peer stores, NCCL all-gather) must reproduce the schedule composed from single-process decoder calls, bit for bit."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("fused", ["1", "0"])
def test_sharded_schedule_two_gpus(fused):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, M3R_FUSED_GATHER=fused)
    port = 29600 + int(fused)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded.py")],
                         env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    assert out.stdout.count("memory == composed: True") == 2
