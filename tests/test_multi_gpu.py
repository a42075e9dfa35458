"""Multi-GPU (needs >= 2 GPUs, skipped otherwise): the sharded evaluation whose payload all-reduce is fused
into the kernel over NVLink peer memory must agree with the NCCL path and leave bit-identical logp on every
rank (tools/peer_check.py asserts both and prints the timing of the two exchanges)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_fused_peer_exchange_matches_nccl(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, M="60000", G="300", C="8", MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), os.path.join(ROOT, "tools", "peer_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "bit-identical on all ranks: True" in out.stdout


@pytest.mark.parametrize("world,model", [(2, "gene"), (4, "gene"), (8, "gene"), (2, "gw")])
def test_sharded_sampler_matches_single_gpu(world, model):
    """Gene-block sharded NUTS: every rank samples its rows of q, the per-chain sums (kinetic
    energy, U-turn dots) are all-reduced over peer memory inside the scalar kernels; all ranks must take
    bit-identical decisions and reproduce the single-GPU chains (tools/sharded_sampler_check.py)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, M="60000", G="300", C="8", TUNE="12", DRAWS="6", MODEL=model, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), os.path.join(ROOT, "tools", "sharded_sampler_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED SAMPLER OK" in out.stdout
