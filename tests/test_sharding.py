"""Multi-rank path.

* CPU, world_size 2 over gloo: shard planning, payload layout and the all-reduce plumbing, with the
  local pass emulated from the oracle (kernels cannot run without a GPU).
* GPU (-m gpu): the real kernels with R ranks emulated one after the other on ONE device — the ranks'
  payloads are summed on the device exactly as the NCCL all-reduce would (B200_PROFILING.md: "emulate
  the ranks ... or test the multi-rank path on the CPU"; never wait across launches on one GPU).
"""
import os
import socket

import numpy as np
import pytest

from baghera_b200 import synth
from baghera_b200.sharding import assemble_global_grad, shard_dataset, shard_snps_only
from oracle import model_oracle as mo
from helpers import REL_TOL, grad_rel_err


def _payload_and_grad_from_oracle(sh, Qloc, c, rank, fix=False):
    """What bgh_logp_grad_local produces for one rank, restated with numpy: payload rows
    {sum r^2/l, sum r/l, sum(beta-mi) [genes whose prior terms this rank counts], sum(beta-mi)^2, residual sums of the
    replicated genes} and the gradient rows of the rank's own (whole) genes."""
    C = Qloc.shape[0]
    chi2, ld = sh.chi2.astype(np.float64), sh.ld.astype(np.float64)
    n_rep = sh.n_replicated
    payload = np.zeros((4 + n_rep, C))
    grad = np.full((C, 3 + sh.n_genes), np.nan)
    for i in range(C):
        q = Qloc[i]
        W, e, mi = np.exp(q[0]), (1.0 if fix else q[1]), 1 / (1 + np.exp(-q[2]))
        beta = q[3:]
        r = chi2 - e - c * ld * beta[sh.gene_id]
        payload[0, i] = np.sum(r * r / ld)
        payload[1, i] = np.sum(r / ld)
        owned = np.ones(sh.n_genes, bool)
        owned[:n_rep] = rank == 0                 # a replicated gene's prior terms are counted by rank 0 only
        d = beta - mi
        payload[2, i] = np.sum(d[owned])
        payload[3, i] = np.sum(d[owned] ** 2)
        sums = np.bincount(sh.gene_id, weights=r, minlength=sh.n_genes)
        g = c * sums - d / W ** 2
        payload[4:, i] = sums[:n_rep]
        g[:n_rep] = np.nan
        grad[i, 3:] = g
    return payload, grad


def _finish_from_payload(sh, Qloc, payload, grad, c, G_total, lik_const_total, fix=False):
    C = Qloc.shape[0]
    lp = np.empty(C)
    for i in range(C):
        q = Qloc[i]
        W, e, mi = np.exp(q[0]), q[1], 1 / (1 + np.exp(-q[2]))
        A_ll, A_e, S1, S2 = payload[0, i], payload[1, i], payload[2, i], payload[3, i]
        lp[i] = ((-1 / W - 2 * np.log(W)) + q[0] - 0.5 * (e - 1) ** 2 - 0.5 * mo.LOG_2PI + np.log(mi) + np.log1p(-mi)
                 - S2 / (2 * W * W) - G_total * np.log(W) - 0.5 * G_total * mo.LOG_2PI - 0.5 * A_ll + lik_const_total)
        grad[i, 0] = 1 / W - 1 + S2 / W ** 2 - G_total
        grad[i, 1] = -(e - 1) + (0.0 if fix else A_e)
        grad[i, 2] = 1 - 2 * mi + mi * (1 - mi) * S1 / W ** 2
        for j in range(sh.n_replicated):
            grad[i, 3 + j] = c * payload[4 + j, i] - (q[3 + j] - mi) / W ** 2
    return lp, grad


@pytest.mark.parametrize("world", [2, 3, 8])
def test_shard_plan(world):
    d = synth.make_arrays(20_000, 200)
    shards = [shard_dataset(d.chi2, d.ld, d.gene_id, d.G, r, world) for r in range(world)]
    assert sum(len(s.chi2) for s in shards) == d.M
    lens = [len(s.chi2) for s in shards]
    assert max(lens) - min(lens) <= 0.2 * d.M / world                   # whole genes, balanced by SNP count (200 genes: coarse)
    counts = np.bincount(d.gene_id, minlength=d.G)
    for r, s in enumerate(shards):
        assert s.gene_id.min() == 0 and s.gene_id.max() == s.n_genes - 1
        assert s.cfg["n_shared_rows"] == s.cfg["n_replicated"] == len(s.shared_genes) == s.n_replicated
        assert s.shared_genes == shards[0].shared_genes                 # every rank holds every replicated gene
        assert s.cfg["owns_replicated"] == int(r == 0)
        # a non-replicated gene lives on exactly one rank, with all of its SNPs
        own = s.gene_ids[s.n_replicated:]
        assert np.array_equal(np.bincount(s.gene_id, minlength=s.n_genes)[s.n_replicated:], counts[own])
    # NonCoding (45% of SNPs, last label) is replicated
    assert d.G - 1 in shards[-1].shared_genes
    allg = np.concatenate([s.gene_ids[s.n_replicated:] for s in shards] + [np.asarray(shards[0].shared_genes, dtype=np.int64)])
    assert np.array_equal(np.sort(allg), np.arange(d.G))


@pytest.mark.parametrize("world", [2, 3, 5, 8])
def test_every_global_row_is_owned_once(world):
    """Sharded sampler: the per-chain sums are exchanged between ranks, so every global coordinate must be counted by
    exactly one rank (replicated head rows and replicated genes included)."""
    d = synth.make_arrays(20_000, 200)
    shards = [shard_dataset(d.chi2, d.ld, d.gene_id, d.G, r, world) for r in range(world)]
    count = np.zeros(3 + d.G, dtype=int)
    for r, sh in enumerate(shards):
        np.add.at(count, sh.row_coords(), sh.owned_rows(r).astype(int))
    assert np.all(count == 1), np.flatnonzero(count != 1)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        d = synth.make_arrays(20_000, 200)
        C = 3
        Q = synth.jittered_start(d.G + 3, C, seed=4)
        sh = shard_dataset(d.chi2, d.ld, d.gene_id, d.G, rank, world)
        Qloc = sh.local_q(Q)
        payload, grad = _payload_and_grad_from_oracle(sh, Qloc, d.c, rank)
        t = torch.from_numpy(payload)
        dist.all_reduce(t)                                            # the per-step exchange
        const = torch.tensor([mo.lik_const(sh.ld.astype(np.float64))], dtype=torch.float64)
        dist.all_reduce(const)                                        # one-off: total normalising constant
        lp, grad = _finish_from_payload(sh, Qloc, t.numpy(), grad, d.c, d.G, float(const[0]))
        q.put((rank, lp, grad, sh))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_matches_unsharded_oracle():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    d = synth.make_arrays(20_000, 200)
    Q = synth.jittered_start(d.G + 3, 3, seed=4)
    lp_ref, g_ref = mo.batch_closed(Q, d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64), d.c)
    grads = assemble_global_grad([r[2] for r in res], [r[3] for r in res], d.G + 3)
    for rank, lp, _, _ in res:
        assert np.allclose(lp, lp_ref, rtol=1e-12)                     # every rank holds the same logp
    for i in range(3):
        assert grad_rel_err(grads[i], g_ref[i]) < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("world,M,G", [(2, 20_000, 200), (4, 60_000, 300), (8, 1_100_000, 15_000)])
def test_emulated_ranks_on_one_gpu(world, M, G):
    import torch
    from baghera_b200.engine import Engine
    d = synth.make_arrays(M, G)
    C = 64 if M > 100_000 else 8
    Q = synth.jittered_start(d.G + 3, C, seed=6)
    shards = [shard_dataset(d.chi2, d.ld, d.gene_id, d.G, r, world) for r in range(world)]
    engines = [Engine(s.chi2, s.ld, s.gene_id, s.n_genes, d.c, C, shard=s.cfg) for s in shards]
    Cp = engines[0].Cp
    rows = engines[0].payload_rows
    qs = [e.to_device_layout(s.local_q(Q)) for e, s in zip(engines, shards)]
    grads = [torch.zeros_like(q) for q in qs]
    payloads = [torch.zeros((rows, Cp), dtype=torch.float64, device="cuda") for _ in range(world)]
    for e, q, g, p in zip(engines, qs, grads, payloads):
        e.logp_grad_local(q, g, p)
    total = torch.stack(payloads).sum(dim=0)                           # what ncclAllReduce(sum) delivers
    # an in-place all-reduce leaves the TOTALS in every rank's payload buffer; the next evaluation must not
    # re-add them (rows of rank-shared genes a rank does not hold have to be zeroed by every launch)
    for p in payloads:
        p.copy_(total)
    for e, q, g, p in zip(engines, qs, grads, payloads):
        e.logp_grad_local(q, g, p)
    total2 = torch.stack(payloads).sum(dim=0)
    assert torch.equal(total, total2), "second evaluation differs: stale totals were re-added to the payload"
    const = sum(e.lik_const for e in engines)                          # one-off all-reduce at setup
    lps = []
    for e, q, g in zip(engines, qs, grads):
        e.set_lik_const(const)
        lp = torch.empty(Cp, dtype=torch.float64, device="cuda")
        e.logp_grad_finish(q, total, lp, g)
        lps.append(lp[:C].cpu().numpy())
    torch.cuda.synchronize()
    n_ref = C if M <= 100_000 else 2
    lp_ref, g_ref = mo.batch_closed(Q[:n_ref], d.chi2.astype(np.float64), d.ld.astype(np.float64),
                                    d.gene_id.astype(np.int64), d.c)
    g_all = assemble_global_grad([e.from_device_layout(g).cpu().numpy() for e, g in zip(engines, grads)], shards, d.G + 3)
    for lp in lps:
        assert np.max(np.abs(lp[:n_ref] - lp_ref) / np.abs(lp_ref)) < REL_TOL
    for i in range(n_ref):
        assert grad_rel_err(g_all[i], g_ref[i]) < REL_TOL
    for e in engines:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world,M,G,model", [(2, 20_000, 200, "gene"), (4, 60_000, 300, "gene"), (8, 1_100_000, 15_000, "gene"),
                                             (3, 30_000, 100, "gamma"), (4, 400_000, 1, "gw")])
def test_fused_exchange_kernel_with_ranks_emulated_on_one_gpu(world, M, G, model):
    """The kernel that ships for N > 1 (bgh_logp_grad_sharded: likelihood pass | grid barrier | per-chain closing with the
    LL-protocol peer exchange) against the oracle on ONE GPU: the ranks' contexts are connected in-process and run one
    after the other — first every rank deposits its payload in the peers' mailboxes without waiting (send-only), then
    every rank runs normally with the same mailbox sequence number and finds the peers' data waiting."""
    import torch
    from baghera_b200.engine import Engine
    d = synth.make_arrays(M, G)
    C = 64 if M > 300_000 and model == "gene" else 8
    gw, gamma = model == "gw", model == "gamma"
    D = 2 if gw else (d.G + 2 if gamma else d.G + 3)
    head = 2 if (gw or gamma) else 3
    c = d.n_patients if gamma else d.c
    Q = synth.jittered_start(D, C, seed=6, model=model)
    if gw:
        shards = [shard_snps_only(d.chi2, d.ld, r, world) for r in range(world)]
    else:
        shards = [shard_dataset(d.chi2, d.ld, d.gene_id, d.G, r, world) for r in range(world)]
    engines = [Engine(s.chi2, s.ld, None if gw else s.gene_id, s.n_genes, c, C, model=model, shard=s.cfg,
                      gamma_rate=d.N_1kG if gamma else None) for s in shards]
    const = sum(e.lik_const for e in engines)
    for e in engines:
        e.set_lik_const(const)
        e.peer_export()
    for r, e in enumerate(engines):
        e.peer_connect_local(r, engines)
    Cp = engines[0].Cp
    qs = [e.to_device_layout(Q if gw else s.local_q(Q, head)) for e, s in zip(engines, shards)]
    grads = [torch.zeros_like(q) for q in qs]
    lps = [torch.zeros(Cp, dtype=torch.float64, device="cuda") for _ in engines]
    for seq in range(2):                                              # twice: the mailbox parities / tags advance
        for e, q, g, lp in zip(engines, qs, grads, lps):
            e.debug_peer_mode(True, seq)
            e.logp_grad_sharded(q, lp, g)
        torch.cuda.synchronize()
        for e, q, g, lp in zip(engines, qs, grads, lps):
            e.debug_peer_mode(False, seq)
            e.logp_grad_sharded(q, lp, g)
        torch.cuda.synchronize()
        assert all(e.peer_status() == 0 for e in engines)
    n_ref = min(C, 2 if M > 300_000 and not gw else 8)
    chi2, ld = d.chi2.astype(np.float64), d.ld.astype(np.float64)
    if gw:
        ref = [mo.gw_logp_grad_closed(Q[i], chi2, ld, d.c) for i in range(n_ref)]
    elif gamma:
        ref = [mo.gamma_logp_grad_closed(Q[i], chi2, ld, d.gene_id.astype(np.int64), float(d.n_patients), float(d.N_1kG)) for i in range(n_ref)]
    else:
        lp_ref, g_ref = mo.batch_closed(Q[:n_ref], chi2, ld, d.gene_id.astype(np.int64), d.c)
        ref = list(zip(lp_ref, g_ref))
    g_loc = [e.from_device_layout(g).cpu().numpy() for e, g in zip(engines, grads)]
    g_all = Q * np.nan if gw else assemble_global_grad(g_loc, shards, D, head)
    for lp in lps[1:]:
        assert torch.equal(lp[:C], lps[0][:C]), "logp differs between ranks (rank-ordered sums must be bit-identical)"
    for i in range(n_ref):
        assert abs(lps[0][i].item() - ref[i][0]) <= REL_TOL * abs(ref[i][0])
        if gw:
            for gl in g_loc:
                assert grad_rel_err(gl[i], ref[i][1]) <= REL_TOL
        else:
            assert grad_rel_err(g_all[i], ref[i][1]) <= REL_TOL
    # replicated rows carry the same gradient on every rank
    nrep = shards[0].n_replicated
    if not gw:
        for gl in g_loc[1:]:
            assert np.array_equal(gl[:, :head + nrep], g_loc[0][:, :head + nrep])
    for e in engines:
        e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 8])
def test_gw_model_emulated_ranks(world):
    """cfg5 shape (genome-wide model, ref: heritability.py:44-55) sharded by SNP ranges: the single slope's
    residual sum is the rank-shared payload row."""
    import torch
    from baghera_b200.engine import Engine
    d = synth.make_arrays(400_000, 1)
    C = 8
    Q = synth.jittered_start(2, C, seed=9, model="gw")
    shards = [shard_snps_only(d.chi2, d.ld, r, world) for r in range(world)]
    engines = [Engine(s.chi2, s.ld, None, 1, d.c, C, model="gw", shard=s.cfg) for s in shards]
    Cp, rows = engines[0].Cp, engines[0].payload_rows
    assert rows == 5
    qs = [e.to_device_layout(Q) for e in engines]
    grads = [torch.zeros_like(q) for q in qs]
    payloads = [torch.zeros((rows, Cp), dtype=torch.float64, device="cuda") for _ in range(world)]
    for e, q, g, p in zip(engines, qs, grads, payloads):
        e.logp_grad_local(q, g, p)
    total = torch.stack(payloads).sum(dim=0)
    const = sum(e.lik_const for e in engines)
    chi2, ld = d.chi2.astype(np.float64), d.ld.astype(np.float64)
    for e, q, g in zip(engines, qs, grads):
        e.set_lik_const(const)
        lp = torch.empty(Cp, dtype=torch.float64, device="cuda")
        e.logp_grad_finish(q, total, lp, g)
        torch.cuda.synchronize()
        gg = e.from_device_layout(g).cpu().numpy()
        for i in range(C):
            lp_ref, g_ref = mo.gw_logp_grad_closed(Q[i], chi2, ld, d.c)
            assert abs(lp[i].item() - lp_ref) <= REL_TOL * abs(lp_ref)
            assert grad_rel_err(gg[i], g_ref) <= REL_TOL
    for e in engines:
        e.close()
