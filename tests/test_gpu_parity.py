"""GPU parity: the CUDA path (through the C ABI) against the oracle on identical inputs.

Bar (BASELINE.json north_star): logp and every gradient component within 1e-6 relative
(fp64 accumulation).  The kernels compute in fp64 throughout, so the observed error is ~1e-13.
"""
import numpy as np
import pytest

from baghera_b200 import synth
from oracle import model_oracle as mo
from helpers import REL_TOL, grad_rel_err

pytestmark = pytest.mark.gpu


def _engine(d, C, **kw):
    from baghera_b200.engine import Engine
    return Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, C, **kw)


def _oracle_batch(d, Q, fix=False):
    return mo.batch_closed(Q, d.chi2.astype(np.float64), d.ld.astype(np.float64),
                           d.gene_id.astype(np.int64), d.c, fix)


def _check(eng, d, Q, fix=False, tol=REL_TOL):
    import torch
    q_dev = eng.to_device_layout(Q)
    logp, grad = eng.logp_grad(q_dev)
    torch.cuda.synchronize()
    lp = logp[:eng.C].cpu().numpy()
    g = eng.from_device_layout(grad).cpu().numpy()
    lp_ref, g_ref = _oracle_batch(d, Q, fix)
    assert np.all(np.isfinite(lp)) and np.all(np.isfinite(g))
    assert np.max(np.abs(lp - lp_ref) / np.abs(lp_ref)) <= tol
    for i in range(Q.shape[0]):
        assert grad_rel_err(g[i], g_ref[i]) <= tol, "chain %d" % i
    return lp, g


@pytest.mark.parametrize("C", [1, 3, 4, 8, 13, 16, 32, 33, 64, 100])
def test_cfg1_all_chain_geometries(cfg1, C):
    """cfg1 (20k SNPs x 200 genes): every chain-lane geometry incl. padded chain counts."""
    eng = _engine(cfg1, C)
    Q = synth.jittered_start(cfg1.G + 3, C, seed=100 + C)
    _check(eng, cfg1, Q)
    eng.close()


def test_fix_intercept(cfg1):
    eng = _engine(cfg1, 4, fix_intercept=True)
    Q = synth.jittered_start(cfg1.G + 3, 4, seed=3)
    _check(eng, cfg1, Q, fix=True)
    eng.close()


def test_edge_cases_small():
    """Gene with one SNP, l == 1 exactly, beta < 0, extreme transforms, M not a multiple of anything."""
    rng = np.random.default_rng(0)
    sizes = np.array([1, 1, 2, 3, 1, 700, 5, 1, 257, 30], dtype=np.int64)
    G, M = len(sizes), int(sizes.sum())
    gid = np.repeat(np.arange(G, dtype=np.int32), sizes)
    ld = (1.0 + np.round(rng.lognormal(3.5, 1.0, M), 3)).astype(np.float32)
    ld[:5] = 1.0
    chi2 = rng.chisquare(1, M).astype(np.float32) * 3
    perm = rng.permutation(M)
    d = synth.SynthData(chi2[perm], ld[perm], gid[perm], ["g%d" % i for i in range(G)], 361194, 1290028, {})
    for C in (4, 64):
        eng = _engine(d, C)
        Q = rng.normal(0.0, 1.0, size=(C, G + 3))
        Q[0, 0], Q[0, 2] = -20.0, 20.0
        Q[1 % C, 0], Q[1 % C, 2] = 20.0, -20.0
        Q[:, 3:] -= 0.5                      # negative betas
        _check(eng, d, Q)
        eng.close()


def test_tiny_inputs():
    """M smaller than the number of lane groups (most ranges empty); single SNP."""
    rng = np.random.default_rng(1)
    for M, G in ((1, 1), (7, 3), (130, 2)):
        gid = np.sort(rng.integers(0, G, M)).astype(np.int32)
        gid[:G] = np.arange(G)
        gid = np.sort(gid)
        d = synth.SynthData(rng.chisquare(1, M).astype(np.float32),
                            (1 + rng.lognormal(3, 1, M)).astype(np.float32), gid,
                            ["g%d" % i for i in range(G)], 361194, 1290028, {})
        eng = _engine(d, 4)
        _check(eng, d, rng.normal(0, 1, size=(4, G + 3)))
        eng.close()


def test_golden_vectors():
    """Committed golden vectors (tests/golden/make_golden.py): q -> (logp, grad) frozen from the oracle."""
    import os
    path = os.path.join(os.path.dirname(__file__), "golden", "logp_grad_cfg1.npz")
    z = np.load(path)
    d = synth.make_arrays(int(z["M"]), int(z["G"]), seed=int(z["seed"]))
    for fix, key in ((False, "free"), (True, "fix")):
        Q = z["Q"]
        eng = _engine(d, Q.shape[0], fix_intercept=fix)
        import torch
        logp, grad = eng.logp_grad(eng.to_device_layout(Q))
        torch.cuda.synchronize()
        lp = logp[:eng.C].cpu().numpy()
        g = eng.from_device_layout(grad).cpu().numpy()
        assert np.max(np.abs(lp - z["logp_" + key]) / np.abs(z["logp_" + key])) <= REL_TOL
        for i in range(Q.shape[0]):
            assert grad_rel_err(g[i], z["grad_" + key][i]) <= REL_TOL
        eng.close()


def test_host_entry_matches_device(cfg1):
    """bgh_logp_grad_host (PyMC3 layout, host buffers, H2D+D2H inside) == device entry == oracle."""
    C = 8
    eng = _engine(cfg1, C)
    Q = synth.jittered_start(cfg1.G + 3, C, seed=5)
    lp_h, g_h = eng.logp_grad_host(Q)
    lp_d, g_d = _check(eng, cfg1, Q)
    assert np.array_equal(lp_h, lp_d) and np.array_equal(g_h, g_d)
    # pipelined multi-eval form: every batch equals its single evaluation
    Q3 = np.stack([Q, Q + 0.01, Q - 0.02])
    lp3, g3 = eng.logp_grad_host(Q3)
    for i in range(3):
        lp1, g1 = eng.logp_grad_host(Q3[i])
        assert np.array_equal(lp3[i], lp1) and np.array_equal(g3[i], g1)
    eng.close()


def test_deterministic_bitwise(cfg1):
    """Fixed summation order: two evaluations are bit-identical; chains do not interact."""
    import torch
    eng = _engine(cfg1, 64)
    Q = synth.jittered_start(cfg1.G + 3, 64, seed=8)
    q_dev = eng.to_device_layout(Q)
    lp1, g1 = eng.logp_grad(q_dev)
    lp2, g2 = eng.logp_grad(q_dev)
    torch.cuda.synchronize()
    assert torch.equal(lp1, lp2) and torch.equal(g1, g2)
    # permuting chains permutes outputs exactly
    perm = np.random.default_rng(0).permutation(64)
    lp3, g3 = eng.logp_grad(eng.to_device_layout(Q[perm]))
    torch.cuda.synchronize()
    assert torch.equal(lp3, lp1[torch.as_tensor(perm, device=lp1.device)])
    eng.close()


def test_gw_model():
    d = synth.make_arrays(200_000, 1, seed=7)
    from baghera_b200.engine import Engine
    import torch
    for fix in (False, True):
        C = 4
        eng = Engine(d.chi2, d.ld, None, 1, d.c, C, model="gw", fix_intercept=fix)
        rng = np.random.default_rng(2)
        Q = np.stack([rng.normal(1.0, 0.5, C), rng.normal(-0.7, 1.0, C)], axis=1)
        logp, grad = eng.logp_grad(eng.to_device_layout(Q))
        torch.cuda.synchronize()
        lp = logp[:C].cpu().numpy()
        g = eng.from_device_layout(grad).cpu().numpy()
        chi2, ld = d.chi2.astype(np.float64), d.ld.astype(np.float64)
        for i in range(C):
            if fix:
                lp_ref, g_ref = mo.gw_fix_logp_grad_closed(Q[i], chi2, ld, d.c)
            else:
                lp_ref, g_ref = mo.gw_logp_grad_closed(Q[i], chi2, ld, d.c)
            assert abs(lp[i] - lp_ref) <= REL_TOL * abs(lp_ref)
            assert grad_rel_err(g[i], g_ref) <= REL_TOL
        eng.close()


def test_full_size_properties():
    """cfg2/cfg3 shape (1.1M x 15k): oracle on 2 chains + size-independent properties on 64."""
    import torch
    d = synth.make_arrays(1_100_000, 15_000)
    C = 64
    eng = _engine(d, C)
    Q = synth.jittered_start(d.G + 3, C, seed=12)
    q_dev = eng.to_device_layout(Q)
    logp, grad = eng.logp_grad(q_dev)
    torch.cuda.synchronize()
    lp = logp[:C].cpu().numpy()
    g = eng.from_device_layout(grad).cpu().numpy()
    lp_ref, g_ref = _oracle_batch(d, Q[[0, 63]])
    for k, i in enumerate((0, 63)):
        assert abs(lp[i] - lp_ref[k]) <= REL_TOL * abs(lp_ref[k])
        assert grad_rel_err(g[i], g_ref[k]) <= REL_TOL
    # property 1: sufficient-statistics identity (an independent O(G) evaluation) for all 64 chains
    ss = mo.SuffStats(d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64), d.G)
    for i in range(0, C, 7):
        lp_s, g_s = mo.logp_grad_suffstat(Q[i], ss, d.c)
        assert abs(lp[i] - lp_s) <= REL_TOL * abs(lp_s) and grad_rel_err(g[i], g_s) <= REL_TOL
    # property 2: the likelihood is quadratic in (e, beta) -> the gradient is affine along a line in beta
    dq = np.zeros_like(Q)
    dq[:, 3:] = np.random.default_rng(3).normal(size=(C, d.G))
    gs = []
    for t in (0.0, 0.5, 1.0):
        _, gt = eng.logp_grad(eng.to_device_layout(Q + t * dq))
        torch.cuda.synchronize()
        gs.append(eng.from_device_layout(gt).cpu().numpy()[:, 3:])
    mid = 0.5 * (gs[0] + gs[2])
    assert np.max(np.abs(gs[1] - mid)) <= 1e-9 * np.max(np.abs(mid))
    eng.close()


@pytest.mark.parametrize("C,fix", [(4, False), (64, False), (8, True)])
def test_gamma_model(cfg1, C, fix):
    """-m gamma (ref: gene_regression.py:229-248): q = [e, mi_logodds__, beta_log__...], slope n_patients * exp(beta_log)."""
    import torch
    from baghera_b200.engine import Engine
    d = cfg1
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, float(d.n_patients), C, model="gamma", fix_intercept=fix,
                 gamma_rate=float(d.N_1kG))
    assert eng.D == d.G + 2
    Q = synth.jittered_start(d.G + 2, C, seed=40 + C, model="gamma")
    Q[:, 0] = 1.0 + 1e-3 * np.random.default_rng(C).normal(size=C)      # e ~ N(1, 0.001)
    logp, grad = eng.logp_grad(eng.to_device_layout(Q))
    torch.cuda.synchronize()
    lp = logp[:C].cpu().numpy()
    g = eng.from_device_layout(grad).cpu().numpy()
    lp_ref, g_ref = mo.gamma_batch_closed(Q, d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64),
                                          float(d.n_patients), float(d.N_1kG), fix)
    assert np.max(np.abs(lp - lp_ref) / np.abs(lp_ref)) <= REL_TOL
    for i in range(C):
        assert grad_rel_err(g[i], g_ref[i]) <= REL_TOL, "chain %d" % i
    eng.close()


def test_cfg2_four_chains_at_full_size():
    """cfg2: 1.1M SNPs x 15k genes with the reference's default 4 chains — the <4,1> chain-lane geometry (8 SNP ranges per
    warp) at full size, every chain against the oracle."""
    d = synth.make_arrays(1_100_000, 15_000)
    eng = _engine(d, 4)
    _check(eng, d, synth.jittered_start(d.G + 3, 4, seed=21))
    eng.close()


def test_cfg5_genome_wide_model_at_ten_million_snps():
    """cfg5: genome-wide model (ref: heritability.py:44-55) on a 10M-SNP synthetic set, 64 chains; oracle on 3 chains plus the
    exact identity sum r^2/l = S3 - 2 e S2 + e^2 S1 - 2 b (B - e n) + b^2 A (sufficient statistics) on all of them."""
    import torch
    from baghera_b200.engine import Engine
    d = synth.make_arrays(10_000_000, 1, seed=11)
    C = 64
    eng = Engine(d.chi2, d.ld, None, 1, d.c, C, model="gw")
    rng = np.random.default_rng(5)
    Q = np.stack([rng.normal(1.0, 0.3, C), rng.normal(-1.0, 1.0, C)], axis=1)
    logp, grad = eng.logp_grad(eng.to_device_layout(Q))
    torch.cuda.synchronize()
    lp = logp[:C].cpu().numpy()
    g = eng.from_device_layout(grad).cpu().numpy()
    chi2, ld = d.chi2.astype(np.float64), d.ld.astype(np.float64)
    for i in (0, 31, 63):
        lp_ref, g_ref = mo.gw_logp_grad_closed(Q[i], chi2, ld, d.c)
        assert abs(lp[i] - lp_ref) <= REL_TOL * abs(lp_ref)
        assert grad_rel_err(g[i], g_ref) <= REL_TOL
    # all chains: d logp / d e = -(e - 1) + S2 - e S1 - c mi A1, with S1 = sum 1/l, S2 = sum s/l, A1 = number of SNPs
    S1, S2, n = np.sum(1.0 / ld), np.sum(chi2 / ld), float(len(ld))
    mi = 1.0 / (1.0 + np.exp(-Q[:, 1]))
    ge = -(Q[:, 0] - 1.0) + S2 - Q[:, 0] * S1 - d.c * mi * n
    assert np.max(np.abs(g[:, 0] - ge) / np.abs(ge)) <= REL_TOL
    eng.close()


def test_gamma_model_at_full_size():
    """-m gamma at 1.1M SNPs x 15k genes, 64 chains: oracle on 2 chains."""
    import torch
    from baghera_b200.engine import Engine
    d = synth.make_arrays(1_100_000, 15_000)
    C = 64
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, float(d.n_patients), C, model="gamma", gamma_rate=float(d.N_1kG))
    Q = synth.jittered_start(d.G + 2, C, seed=77, model="gamma")
    Q[:, 0] = 1.0 + 1e-3 * np.random.default_rng(1).normal(size=C)
    logp, grad = eng.logp_grad(eng.to_device_layout(Q))
    torch.cuda.synchronize()
    lp = logp[:C].cpu().numpy()
    g = eng.from_device_layout(grad).cpu().numpy()
    lp_ref, g_ref = mo.gamma_batch_closed(Q[[0, 63]], d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64),
                                          float(d.n_patients), float(d.N_1kG), False)
    for k, i in enumerate((0, 63)):
        assert abs(lp[i] - lp_ref[k]) <= REL_TOL * abs(lp_ref[k])
        assert grad_rel_err(g[i], g_ref[k]) <= REL_TOL
    eng.close()


@pytest.mark.parametrize("M,G", [(20_000, 200), (1_100_000, 15_000), (37, 5)])
def test_device_built_stream_is_the_stable_gene_sort(M, G):
    """bgh_upload builds the sorted / padded / prepared fp64 stream on the device (prepare_stream_kernel): every bit
    must equal the host formulas on a STABLE sort by gene id (file order kept inside a gene, ref: Q6), every gene
    padded to a multiple of 8 zero-weight SNPs, and the likelihood constant -0.5 sum log(2 pi l)."""
    d = synth.make_arrays(M, G, seed=5)
    rng = np.random.default_rng(0)
    perm = rng.permutation(M)                      # file order is not gene order
    chi2, ld, gid = d.chi2[perm], d.ld[perm], d.gene_id[perm]
    from baghera_b200.engine import Engine
    eng = Engine(chi2, ld, gid, d.G, d.c, 4)
    sp, wp, lp, g_dev = eng.device_stream()
    order = np.argsort(gid, kind="stable")
    cnt = np.bincount(gid, minlength=d.G)
    pcnt = np.concatenate([[0], np.cumsum((cnt + 7) // 8 * 8)])
    start = np.concatenate([[0], np.cumsum(cnt)])[:-1]
    gs = gid[order]
    dest = pcnt[gs] + (np.arange(M) - start[gs])
    l = ld[order].astype(np.float64)
    sq = np.sqrt(l)
    ref = np.zeros((3, sp.size))
    ref[0, dest] = chi2[order].astype(np.float64) / sq
    ref[1, dest] = 1.0 / sq
    ref[2, dest] = sq
    assert np.array_equal(sp, ref[0]) and np.array_equal(wp, ref[1]) and np.array_equal(lp, ref[2])
    Mp = int(pcnt[-1])
    assert np.array_equal(g_dev[:Mp], np.repeat(np.arange(d.G), np.diff(pcnt)))
    ref_const = -0.5 * np.sum(np.log(ld.astype(np.float64))) - 0.5 * M * np.log(2 * np.pi)
    assert abs(eng.lik_const - ref_const) <= 1e-13 * abs(ref_const)
    eng.close()


def test_upload_rejects_bad_ld():
    """ld <= 0 or non-finite inputs are refused by the device-side validation (BGH_EINVAL), not silently used."""
    from baghera_b200.engine import Engine
    from baghera_b200._lib import BghError
    d = synth.make_arrays(2_000, 20, seed=6)
    ld = d.ld.copy()
    ld[1234] = 0.0
    with pytest.raises(BghError):
        Engine(d.chi2, ld, d.gene_id, d.G, d.c, 4)
    chi2 = d.chi2.copy()
    chi2[7] = np.nan
    with pytest.raises(BghError):
        Engine(chi2, d.ld, d.gene_id, d.G, d.c, 4)
