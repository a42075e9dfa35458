"""Gene-block sharding of the gene-sorted SNP array across ranks.

Genes are atomic: every rank owns a contiguous block of whole genes, balanced by SNP count, together with their rows of q,
momentum, mass matrix and trace statistics.  A gene too large for that (NonCoding spans ~45% of all SNPs) is REPLICATED:
every rank holds an equal share of its SNPs and a copy of its row of q; its residual sum travels in the per-step payload
(row 4 + j for replicated gene j) and every rank updates its copy identically.  The shared parameters (rows 0..2 of q) are
replicated the same way.  Per step the only exchange is one all-reduce (sum) of [4 + n_replicated] fp64 per chain for an
evaluation, [28 + n_replicated] for a sampler tick (the sampler's per-chain sums ride in the same message).

Local layout of a shard: rows [0, head) shared parameters, then the replicated genes (local gene ids 0 .. n_rep-1, in
global id order), then the block's genes in global id order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    chi2: np.ndarray          # float32, this rank's SNPs (gene sorted, replicated genes first)
    ld: np.ndarray
    gene_id: np.ndarray       # int32 local ids 0..n_genes-1
    n_genes: int
    gene_ids: np.ndarray      # global id of every local gene (replicated genes first)
    n_replicated: int
    cfg: dict                 # bgh_cfg sharding fields
    shared_genes: list        # global ids of the replicated genes (payload row 4 + k <-> shared_genes[k])

    def row_coords(self, head_rows=3):
        """Global coordinate of every local row of q (head rows keep their index)."""
        return np.concatenate([np.arange(head_rows), head_rows + self.gene_ids]).astype(np.int32)

    def owned_rows(self, rank, head_rows=3):
        """Mask over this rank's local rows of q: True where the rank counts the row in the sampler's per-chain sums
        (kinetic energy, U-turn dots): replicated rows (shared parameters, replicated genes) belong to rank 0."""
        own = np.ones(head_rows + self.n_genes, dtype=bool)
        own[:head_rows + self.n_replicated] = rank == 0
        return own

    def local_q(self, Q, head_rows=3):
        """Global chain-major Q[C, head + G] -> this rank's Q[C, head + n_genes]."""
        return np.ascontiguousarray(Q[:, self.row_coords(head_rows)])


def shard_dataset(chi2, ld, gene_id, n_genes_total, rank, world, replicate_above=0.5):
    """Shard `rank` of `world`.  A gene with more than replicate_above * M / world SNPs is replicated."""
    gene_id = np.asarray(gene_id)
    M = gene_id.shape[0]
    order = np.argsort(gene_id, kind="stable")
    gid = gene_id[order]
    counts = np.bincount(gid, minlength=int(n_genes_total)).astype(np.int64)
    rep = np.flatnonzero(counts > replicate_above * M / world) if world > 1 else np.zeros(0, dtype=np.int64)
    is_rep = np.zeros(int(n_genes_total), dtype=bool)
    is_rep[rep] = True
    starts = np.concatenate([[0], np.cumsum(counts)])
    # block genes: contiguous runs of the non-replicated genes, cut where the cumulative SNP count crosses k / world
    blk_genes = np.flatnonzero(~is_rep & (counts > 0))
    cum = np.cumsum(counts[blk_genes])
    total = int(cum[-1]) if len(cum) else 0
    cuts = []
    for k in range(world + 1):                         # the gene boundary nearest to k / world of the SNPs
        x = k * total / world
        j = int(np.searchsorted(cum, x, side="left"))
        if j < len(cum) and (j == 0 or abs(cum[j] - x) <= abs(cum[j - 1] - x)):
            j += 1
        cuts.append(min(j, len(blk_genes)))
    cuts[0], cuts[-1] = 0, len(blk_genes)
    cuts = list(np.maximum.accumulate(cuts))
    mine = blk_genes[cuts[rank]:cuts[rank + 1]]
    sel, loc = [], []
    for j, g in enumerate(rep):                       # an equal contiguous share of every replicated gene
        a, b = int(starts[g]), int(starts[g + 1])
        lo, hi = a + (b - a) * rank // world, a + (b - a) * (rank + 1) // world
        sel.append(order[lo:hi])
        loc.append(np.full(hi - lo, j, dtype=np.int32))
    for k, g in enumerate(mine):
        a, b = int(starts[g]), int(starts[g + 1])
        sel.append(order[a:b])
        loc.append(np.full(b - a, len(rep) + k, dtype=np.int32))
    if not sel or sum(len(x) for x in sel) == 0:
        raise ValueError("rank %d of %d would own no SNPs (M=%d)" % (rank, world, M))
    sel = np.concatenate(sel)
    loc = np.concatenate(loc)
    if np.bincount(loc, minlength=len(rep) + len(mine)).min() == 0:
        raise ValueError("rank %d of %d: a replicated gene has fewer SNPs than ranks" % (rank, world))
    n_genes = len(rep) + len(mine)
    cfg = dict(n_genes_total=int(n_genes_total), n_shared_rows=len(rep), n_replicated=len(rep), owns_replicated=int(rank == 0),
               first_gene_partial=0, last_gene_partial=0, first_gene_row=-1, last_gene_row=-1)
    return Shard(np.asarray(chi2)[sel], np.asarray(ld)[sel], loc, n_genes, np.concatenate([rep, mine]).astype(np.int64), len(rep),
                 cfg, [int(g) for g in rep])


def shard_snps_only(chi2, ld, rank, world):
    """Genome-wide model (no genes): SNPs sharded evenly; the single slope is the one replicated 'gene'."""
    M = len(chi2)
    lo, hi = M * rank // world, M * (rank + 1) // world
    cfg = dict(n_genes_total=1, n_shared_rows=1, n_replicated=1, owns_replicated=int(rank == 0),
               first_gene_partial=0, last_gene_partial=0, first_gene_row=-1, last_gene_row=-1)
    return Shard(np.asarray(chi2)[lo:hi], np.asarray(ld)[lo:hi], np.zeros(hi - lo, dtype=np.int32), 1, np.zeros(1, dtype=np.int64), 1, cfg, [0])


def assemble_global_grad(shards_grads, shards, D, head_rows=3):
    """Stitches per-rank gradients [C, head + n_genes_r] into the global [C, D] (replicated rows must agree)."""
    C = shards_grads[0].shape[0]
    out = np.full((C, D), np.nan)
    for g, sh in zip(shards_grads, shards):
        out[:, sh.row_coords(head_rows)] = g
    return out


def connect_peers(engine, dist, device):
    """One-off setup of the fused exchange: all-gather the ranks' mailbox IPC handles and map the peers.
    `dist` is an initialised torch.distributed (NCCL or gloo) with one process per GPU of ONE node."""
    import torch

    rank, world = dist.get_rank(), dist.get_world_size()
    h = engine.peer_export()
    backend = dist.get_backend()
    dev = device if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(h), dtype=torch.uint8, device=dev)
    allh = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allh, mine)
    engine.peer_connect(rank, world, [bytes(t.cpu().tolist()) for t in allh])
    dist.barrier()      # every mailbox is mapped (and zeroed) before anybody's first evaluation
