"""Gene-block sharding of the gene-sorted SNP array across ranks.

Each rank owns a contiguous SNP range balanced by SNP count; a gene cut by a rank boundary (NonCoding
spans ~45% of all SNPs) is replicated on the ranks it touches and its residual sum travels in the
all-reduce payload (rows 4..).  Shared parameters (rows 0..2 of q) are replicated.  Per step the only
exchange is one all-reduce (sum) of the payload [4 + n_shared][Cp] fp64.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Shard:
    chi2: np.ndarray          # float32, this rank's SNPs (gene sorted)
    ld: np.ndarray
    gene_id: np.ndarray       # int32 local ids 0..n_genes-1
    n_genes: int
    gene_lo: int              # global id of local gene 0
    cfg: dict                 # bgh_cfg sharding fields
    shared_genes: list        # global ids of the rank-shared genes (payload row k <-> shared_genes[k])

    def owned_rows(self, rank, head_rows=3):
        """Mask over this rank's local rows of q: True where the rank counts the row in the sampler's per-chain
        sums (kinetic energy, U-turn dots).  The rule of nuts_params() in csrc/bgh_api.cu: the shared head rows
        belong to rank 0, a gene that continues from the previous rank belongs to the LOWEST rank holding it."""
        own = np.ones(head_rows + self.n_genes, dtype=bool)
        own[:head_rows] = rank == 0
        if self.cfg["first_gene_partial"]:
            own[head_rows] = False
        return own

    def local_q(self, Q):
        """Global chain-major Q[C, 3 + G] -> this rank's Q[C, 3 + n_genes]."""
        return np.concatenate([Q[:, :3], Q[:, 3 + self.gene_lo: 3 + self.gene_lo + self.n_genes]], axis=1)


def shard_dataset(chi2, ld, gene_id, n_genes_total, rank, world):
    order = np.argsort(gene_id, kind="stable")
    gid = np.asarray(gene_id)[order]
    M = gid.shape[0]
    cuts = [int(round(k * M / world)) for k in range(world + 1)]
    shared = []
    for k in range(1, world):
        x = cuts[k]
        if 0 < x < M and gid[x - 1] == gid[x]:
            g = int(gid[x])
            if not shared or shared[-1] != g:
                shared.append(g)
    a, b = cuts[rank], cuts[rank + 1]
    if b <= a:
        raise ValueError("rank %d of %d would own no SNPs (M=%d)" % (rank, world, M))
    g_lo, g_hi = int(gid[a]), int(gid[b - 1])
    first_partial = a > 0 and gid[a - 1] == gid[a]
    last_partial = b < M and gid[b] == gid[b - 1]
    cfg = dict(n_genes_total=int(n_genes_total), n_shared_rows=len(shared),
               first_gene_partial=int(first_partial), last_gene_partial=int(last_partial),
               first_gene_row=shared.index(g_lo) if g_lo in shared else -1,
               last_gene_row=shared.index(g_hi) if g_hi in shared else -1)
    sel = order[a:b]
    return Shard(np.asarray(chi2)[sel], np.asarray(ld)[sel], (gid[a:b] - g_lo).astype(np.int32), g_hi - g_lo + 1, g_lo,
                 cfg, shared)


def assemble_global_grad(shards_grads, shards, D):
    """Stitches per-rank gradients [C, 3 + n_genes_r] into the global [C, D] (replicated rows must agree)."""
    C = shards_grads[0].shape[0]
    out = np.full((C, D), np.nan)
    for g, sh in zip(shards_grads, shards):
        out[:, :3] = g[:, :3]
        out[:, 3 + sh.gene_lo: 3 + sh.gene_lo + sh.n_genes] = g[:, 3:]
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
