"""Synthetic BAGHERA-shaped inputs.

Two products:

* :func:`make_arrays` – the arrays the hot path consumes *after* the reference's
  data preparation (ref: baghera_tool/snps.py:66-78 ``stats = z**2``;
  ref: baghera_tool/snps.py:117-134 ``l = 1 + l*(l>0)``; ref:
  baghera_tool/gene_regression.py:61-71 gene index = rank of the gene label in
  sorted label order).  Used by bench.py and the parity tests.
* :func:`make_snp_table` – an annotated-SNP table with the on-disk schema of
  ``generate-snp-file`` (ref: docs/createdata.rst:45-47: ``chr, rs_id, position,
  cm, maf, l, gene, sample_size, z``) for the host-layer / CLI tests.

The chi2 / LD values are rounded to fp32 (that is what the device stores); callers
that need fp64 (the oracle) upcast the *rounded* values so that both sides see
identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

N_PATIENTS = 361194          # ref: docs/createdata.rst:47
N_1KG = 1290028              # ref: baghera_tool/command.py:32
SEED = 20211

SHAPES = {
    "cfg1": dict(M=20_000, G=200),
    "cfg2": dict(M=1_100_000, G=15_000),
    "cfg3": dict(M=1_100_000, G=15_000),
    "cfg5": dict(M=10_000_000, G=1),
}


@dataclass
class SynthData:
    """Gene-regression inputs in *file order* (not gene sorted, quirk Q6)."""
    chi2: np.ndarray        # float32[M]   s_j = z_j**2
    ld: np.ndarray          # float32[M]   l_j >= 1
    gene_id: np.ndarray     # int32[M]     rank of the gene label, unsorted
    gene_names: list        # len G, sorted label order
    n_patients: int
    N_1kG: int
    truth: dict

    @property
    def M(self):
        return int(self.chi2.shape[0])

    @property
    def G(self):
        return len(self.gene_names)

    @property
    def c(self):
        return float(self.n_patients) / float(self.N_1kG)


def gene_sizes(M: int, G: int, rng: np.random.Generator, noncoding_frac: float = 0.45):
    """NonCoding gets 45% of M, the other G-1 genes sizes ∝ LogNormal(0,1), min 1."""
    if G == 1:
        return np.array([M], dtype=np.int64)
    n_nc = int(round(noncoding_frac * M))
    rest = M - n_nc
    if n_nc < 1 or rest < G - 1:
        raise ValueError("gene_sizes: M=%d is too small for G=%d genes of at least one SNP" % (M, G))
    w = rng.lognormal(0.0, 1.0, size=G - 1)
    sizes = np.maximum(1, np.floor(w / w.sum() * rest).astype(np.int64))
    # fix the rounding drift on the largest coding genes
    drift = rest - int(sizes.sum())
    order = np.argsort(-sizes)
    i = 0
    while drift != 0:
        k = order[i % (G - 1)]
        if drift > 0:
            sizes[k] += 1
            drift -= 1
        elif sizes[k] > 1:
            sizes[k] -= 1
            drift += 1
        i += 1
    return np.concatenate([sizes, [n_nc]]).astype(np.int64)   # NonCoding sorts last


def make_arrays(M: int, G: int, seed: int = SEED, shuffle: bool = True,
                model_consistent: bool = True) -> SynthData:
    rng = np.random.default_rng(seed)
    e0, mi0, W0 = 1.02, 0.20, 0.15
    sizes = gene_sizes(M, G, rng)
    if G > 1:
        width = len(str(G))
        names = ["G%0*d" % (width, i) for i in range(G - 1)] + ["NonCoding"]
    else:
        names = ["NonCoding"]
    assert names == sorted(names)
    gene_id = np.repeat(np.arange(G, dtype=np.int32), sizes)
    ld = (1.0 + np.round(rng.lognormal(3.5, 1.0, size=M), 3)).astype(np.float32)
    beta = rng.normal(mi0, W0, size=G)
    if G > 1:
        hot = rng.random(G) < 0.01
        hot[-1] = False
        beta[hot] = mi0 + 1.0
    else:
        beta[:] = mi0
    c = N_PATIENTS / N_1KG
    l64 = ld.astype(np.float64)
    if model_consistent:
        chi2 = e0 + c * l64 * beta[gene_id] + np.sqrt(l64) * rng.standard_normal(M)
    else:
        z = rng.standard_normal(M) * np.sqrt(1.0 + c * l64 * np.maximum(beta[gene_id], 0.0))
        chi2 = z * z
    chi2 = chi2.astype(np.float32)
    if shuffle:
        perm = rng.permutation(M)
        chi2, ld, gene_id = chi2[perm], ld[perm], gene_id[perm]
    return SynthData(chi2=chi2, ld=ld, gene_id=gene_id, gene_names=names,
                     n_patients=N_PATIENTS, N_1kG=N_1KG,
                     truth=dict(e=e0, mi=mi0, W=W0, beta=beta))


def make_snp_table(M: int, G: int, seed: int = SEED, neg_ld_frac: float = 0.002,
                   low_maf_frac: float = 0.01):
    """Annotated SNP table as a pandas DataFrame (schema of ref: docs/createdata.rst:45-47).

    ``l`` is the *raw* LD score (the filter adds 1 and clamps negatives, Q7); ``z`` is signed
    with ``E[z**2]`` following the model's mean; NonCoding SNPs carry ``gene = NaN``
    (ref: baghera_tool/snps.py:98-102).
    """
    import pandas as pd

    rng = np.random.default_rng(seed)
    e0, mi0, W0 = 1.02, 0.20, 0.15
    sizes = gene_sizes(M, G, rng)
    width = len(str(G))
    names = ["G%0*d" % (width, i) for i in range(G - 1)] + [None]
    gidx = np.repeat(np.arange(G), sizes)
    raw_l = np.round(rng.lognormal(3.5, 1.0, size=M), 3)
    neg = rng.random(M) < neg_ld_frac
    raw_l[neg] = -np.round(rng.random(int(neg.sum())), 3)
    l_eff = 1.0 + raw_l * (raw_l > 0)
    beta = rng.normal(mi0, W0, size=G)
    hot = rng.random(G) < 0.01
    hot[-1] = False
    beta[hot] = mi0 + 1.0
    c = N_PATIENTS / N_1KG
    # z ~ N(0, sd^2) with sd^2 = e0 + c*l*beta, so that E[z^2] follows the model's mean structure
    # (chi2 statistics cannot be negative, unlike draws from the Normal likelihood itself)
    z = rng.standard_normal(M) * np.sqrt(np.maximum(e0 + c * l_eff * beta[gidx], 0.05))
    maf = rng.uniform(0.011, 0.5, size=M)
    low = rng.random(M) < low_maf_frac
    maf[low] = rng.uniform(0.001, 0.0099, size=int(low.sum()))
    gene_chr = rng.integers(1, 23, size=G)
    chrom = gene_chr[gidx]
    nc = gidx == G - 1
    chrom[nc] = rng.integers(1, 23, size=int(nc.sum()))
    df = pd.DataFrame({
        "chr": chrom.astype(np.int64),
        "rs_id": ["rs%d" % (i + 1) for i in range(M)],
        "position": rng.integers(1, 240_000_000, size=M),
        "cm": np.round(rng.uniform(0, 250, size=M), 3),
        "maf": np.round(maf, 4),
        "l": raw_l,
        "gene": pd.Series([names[g] for g in gidx], dtype=object),
        "sample_size": N_PATIENTS,
        "z": z,
    })
    perm = rng.permutation(M)
    df = df.iloc[perm].reset_index(drop=True)
    return df, dict(e=e0, mi=mi0, W=W0, beta=beta, names=names)


def jittered_start(D: int, C: int, seed: int = 1, model: str = "gene") -> np.ndarray:
    """PyMC3 ``jitter+adapt_diag`` start: test point + U(-1,1) per coordinate, per chain.

    Test point [PyMC3-3.6, recalled]: W -> mode of InverseGamma(1,1) = 0.5 (log-transformed),
    e -> 1, mi -> 0.5 (logodds 0), beta -> 0.5.  Returns float64[C, D] (chain-major).
    """
    rng = np.random.default_rng(seed)
    if model == "gene":
        tp = np.full(D, 0.5)
        tp[0] = np.log(0.5)
        tp[1] = 1.0
        tp[2] = 0.0
    elif model == "gamma":   # [e, mi_logodds__, beta_log__...]; Gamma(mi, N_1kG) test value = mean = 0.5 / N_1kG
        tp = np.full(D, np.log(0.5 / N_1KG))
        tp[0] = 1.0
        tp[1] = 0.0
    else:  # gw: [e, mi_logodds__]; Beta(1,2) mean = 1/3
        tp = np.array([1.0, np.log((1 / 3) / (2 / 3))])
    return tp[None, :] + rng.uniform(-1.0, 1.0, size=(C, D))
