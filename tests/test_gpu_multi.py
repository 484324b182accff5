"""Multi-GPU parity (needs >= 2 GPUs): the sharded engine under torchrun equals the single-GPU engine."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("mode,launch", [("fp32", "eager"), ("bf16", "eager"), ("torus_fp32", "eager"), ("torus_bf16", "eager")])
def test_sharded_engine_matches_single_gpu(mode, launch):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29617", os.path.join(ROOT, "tests", "multi_gpu_check.py"), mode, launch]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert "MULTI_GPU_CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
