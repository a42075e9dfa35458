"""CPU tests of the host-side work plan (bgh_debug_plan) against an emulation of the kernel's rule."""
import numpy as np
import pytest

from baghera_b200 import synth
from baghera_b200.engine import debug_plan
from helpers import emulate_plan_gene_sums


def _sorted_case(M, G, seed=1):
    if M == G:   # one SNP per gene
        rng = np.random.default_rng(seed)
        return np.arange(G, dtype=np.int32), rng.normal(size=M)
    d = synth.make_arrays(M, G, seed=seed)
    order = np.argsort(d.gene_id, kind="stable")
    return d.gene_id[order].astype(np.int32), d.chi2[order].astype(np.float64)


@pytest.mark.parametrize("M,G,n_cta,gpc", [(20_000, 200, 4, 16), (20_000, 200, 148, 16), (5_000, 1, 19, 16),
                                           (300, 7, 64, 16), (1, 1, 1, 16), (4097, 4097, 32, 16),
                                           (100_000, 1500, 148, 128), (50_000, 1, 148, 16), (60_000, 30, 7, 32)])
def test_plan_covers_and_reassembles(built_lib, M, G, n_cta, gpc):
    gid, val = _sorted_case(M, G)
    plan = debug_plan(gid, G, n_cta, gpc)
    off = plan["off"]
    assert off[0] == 0 and off[-1] == M and np.all(np.diff(off) >= 0)
    sums, owners = emulate_plan_gene_sums(plan, gid, val, G)
    ref = np.bincount(gid, weights=val, minlength=G)
    assert np.allclose(sums, ref, rtol=1e-12, atol=1e-9)
    assert np.all(owners == 1), "every gene's prior must be owned by exactly one group"


def test_plan_balance_and_snapping(built_lib):
    gid, _ = _sorted_case(1_100_000, 15_000)
    NG = 148 * 16
    plan = debug_plan(gid, 15_000, 148, 16)
    lens = np.diff(plan["off"])
    ideal = 1_100_000 / NG
    # the plan balances cost = SNPs + kappa * gene boundaries (kappa = 5 SNP-equivalents per flush, plan_kappa in bgh_api.cu)
    starts = np.flatnonzero(np.diff(gid, prepend=-1))
    n_starts = np.diff(np.searchsorted(starts, plan["off"]))
    cost = lens + 5.0 * n_starts
    excl_ = (plan["flags"] & 4) != 0
    cost[excl_] = lens[excl_]
    assert cost.max() <= 1.15 * cost.mean() and cost.min() >= 0.85 * cost.mean()
    per_cta = cost.reshape(148, 16).sum(axis=1)
    assert per_cta.max() <= 1.03 * per_cta.mean()
    # a cut is snapped to a gene boundary within ideal/16, so only genes longer than ideal/8 are split
    counts = np.bincount(gid)
    assert all(counts[g] > 2 * int(0.8 * ideal / 16) for g in plan["split_gene"])
    assert len(plan["split_gene"]) < NG // 4
    # NonCoding (45% of SNPs) gets exclusive CTAs: one slot per CTA, listed first (long list)
    n0 = plan["split_ptr"][1] - plan["split_ptr"][0]
    assert plan["split_gene"][0] == 14_999 and 55 <= n0 <= 70
    excl = (plan["flags"] & 4) != 0
    assert excl.sum() == n0 * 16 and np.all(gid[plan["off"][:-1][excl]] == 14_999)


def test_plan_rank_partial_flags(built_lib):
    """Sharded ranks: first/last gene continue on neighbours -> they must be assembled from slots."""
    gid, val = _sorted_case(20_000, 50)
    plan = debug_plan(gid, 50, 6, 16, first_partial=True, last_partial=True)
    assert 0 in plan["split_gene"] and 49 in plan["split_gene"]
    sums, owners = emulate_plan_gene_sums(plan, gid, val, 50)
    assert np.allclose(sums, np.bincount(gid, weights=val, minlength=50))
    assert owners[0] == 0 and owners[49] == 1 and np.all(owners[1:49] == 1)


def test_plan_rejects_bad_input(built_lib):
    from baghera_b200._lib import BghError
    with pytest.raises(BghError):
        debug_plan(np.array([0, 2, 1], np.int32), 3, 2, 4)      # unsorted
    with pytest.raises(BghError):
        debug_plan(np.array([0, 0, 2], np.int32), 3, 2, 4)      # gene 1 missing
    with pytest.raises(BghError):
        debug_plan(np.array([0, 5], np.int32), 3, 2, 4)         # id out of range


def test_sampler_plan_balances_the_warps_of_a_cta_by_likelihood_cost(built_lib):
    """Sampler plan (bgh_tick2.cu): CTA cuts weigh a gene by the leapfrog of its row (kappa 60), but the cuts between the
    WARPS of a CTA must balance what the likelihood pass costs a warp (SNPs + ~8 per gene boundary): the CTA waits for its
    slowest warp.  With one kappa for both levels a warp holding one long gene streamed ~3x the mean."""
    from baghera_b200.engine import debug_sampler_plan
    d = synth.make_arrays(550_000, 7_500, seed=3)            # the shape of a rank of a 2-way gene-block sharding
    cnt = np.bincount(d.gene_id, minlength=d.G)
    gid = np.repeat(np.arange(d.G, dtype=np.int32), (cnt + 7) // 8 * 8)      # padded to the batch, as on the device
    res = {}
    for two_level in (True, False):
        off, flags = debug_sampler_plan(gid, d.G, 148, 16, two_level)
        assert off[0] == 0 and off[-1] == gid.shape[0] and np.all(np.diff(off) >= 0) and np.all(off[1:-1] % 8 == 0)
        worst = []
        for c in range(148):
            if flags[c * 16] & 4:          # exclusive CTA (one huge gene): equal cuts
                continue
            cost = []
            for v in range(c * 16, c * 16 + 16):
                a, b = off[v], off[v + 1]
                cost.append((b - a) + 8.0 * (len(np.unique(gid[a:b])) if b > a else 0))
            cost = np.array(cost)
            worst.append(cost.max() / cost.mean())
        res[two_level] = (np.median(worst), np.max(worst))
    assert res[True][0] < 1.45 and res[True][0] < 0.9 * res[False][0] and res[True][1] < 0.85 * res[False][1], res
