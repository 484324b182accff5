"""Multi-GPU parity (needs >= 2 GPUs): the vertex-sharded engine under torchrun equals the single-GPU engine, for
every power-of-two world size the box offers (2, 4, 8), eager and under CUDA-graph replay."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worlds():
    """World sizes to test: EP_TEST_WORLDS (comma separated) or, by default, 2 and the largest power of two the box has."""
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    env = os.environ.get("EP_TEST_WORLDS")
    if env:
        return [int(w) for w in env.split(",") if int(w) <= n]
    fit = [w for w in (2, 4, 8) if w <= n]
    return sorted(set(fit[:1] + fit[-1:]))


@pytest.mark.parametrize("mode,launch", [("fp32", "eager"), ("bf16", "eager"), ("bf16", "graph"), ("torus_fp32", "eager"),
                                         ("torus_bf16", "eager"), ("cabi", "eager"), ("cabi", "graph")])
def test_sharded_engine_matches_single_gpu(mode, launch):
    worlds = _worlds()
    if not worlds:
        pytest.skip("needs at least 2 GPUs")
    for world in worlds:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(29617 + world),
               os.path.join(ROOT, "tests", "multi_gpu_check.py"), mode, launch]
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert "MULTI_GPU_CHECK OK" in out.stdout, "world %d\n" % world + out.stdout[-2000:] + out.stderr[-3000:]
