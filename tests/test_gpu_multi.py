"""Multi-GPU checks that need real peer access (skipped on a single-GPU box)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fused_peer_memory_allreduce_adam_matches_nccl():
    """pinn_adam_step_p2p (gradient sum over NVLink peer memory fused into Adam) keeps two replicas bit-identical and
    agrees with the NCCL all-reduce + Adam path (tests/multi_gpu_p2p_check.py under torchrun)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29591", os.path.join(ROOT, "tests", "multi_gpu_p2p_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert "MULTI_GPU_P2P_CHECK" in res.stdout
