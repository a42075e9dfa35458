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


@pytest.mark.parametrize("world,model,driver", [(2, "gene", "persistent"), (4, "gene", "persistent"), (8, "gene", "persistent"),
                                                (2, "gw", "persistent"), (2, "gene", "lockstep")])
def test_sharded_sampler_matches_single_gpu(world, model, driver):
    """Gene-block sharded NUTS: every rank samples its rows of q, the per-chain sums (likelihood sums,
    kinetic energy, U-turn dots) travel in ONE peer exchange per leapfrog inside the persistent kernel (or, lock-step
    driver, are all-reduced inside the scalar kernels); all ranks must take bit-identical decisions and reproduce the
    single-GPU chains (tools/sharded_sampler_check.py)."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, M="60000", G="300", C="8", TUNE="12", DRAWS="6", MODEL=model, DRIVER=driver, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(29560 + world), os.path.join(ROOT, "tools", "sharded_sampler_check.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "SHARDED SAMPLER OK" in out.stdout


def test_gene_heritability_on_two_devices(tmp_path):
    """`--devices 0,1`: the chains of one analysis run as independent replicas over two GPUs (replicas.py); the
    artefacts must have the same shape as the single-device run and agree with it within Monte-Carlo error."""
    import numpy as np
    import pandas as pd
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from baghera_b200 import gene_regression as gr
    from baghera_b200 import synth
    from baghera_b200.cli import main
    df, truth = synth.make_snp_table(12000, 30, seed=8)
    path = str(tmp_path / "snps.csv")
    df.to_csv(path, index=False)
    runs = {}
    for tag, extra in (("one", []), ("two", ["--devices", "0,1"])):
        out = [str(tmp_path / ("%s_%s" % (tag, n))) for n in ("genes.csv", "summary.csv", "log.txt")]
        rc = main(["gene-heritability", path, *out, "--sweeps", "500", "-b", "500", "--n-chains", "8", "-m", "normal",
                   "--seed", "3", *extra])
        assert rc == 0
        runs[tag] = (pd.read_csv(out[0]), pd.read_csv(out[1], index_col=0), dict(gr.analyse_normal.last_run))
    assert runs["one"][2]["devices"] == [0] and runs["two"][2]["devices"] == [0, 1]
    assert runs["two"][2]["stats"].shape == runs["one"][2]["stats"].shape           # [tune + draws, 8, 8 chains]
    g1, s1, _ = runs["one"]
    g2, s2, _ = runs["two"]
    assert list(g1.columns) == list(g2.columns) and set(g1["name"]) == set(g2["name"])
    assert s2["Rhat"].max() < 1.05                                                 # chains from both GPUs mix
    for row in ("mi", "e", "W", "herTOT"):
        tol = 5.0 * max(s1.loc[row, "mc_error"], s2.loc[row, "mc_error"]) + 1e-3 * abs(s1.loc[row, "mean"])
        assert abs(s1.loc[row, "mean"] - s2.loc[row, "mean"]) < tol, (row, s1.loc[row, "mean"], s2.loc[row, "mean"], tol)
    m = g1.merge(g2, on="name", suffixes=("_1", "_2"))
    assert np.corrcoef(m["bg_mean_1"], m["bg_mean_2"])[0, 1] > 0.98
