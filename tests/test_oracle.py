"""CPU tests of the oracle: the four restatements of the reference density agree (SURVEY §8c pins 1-4)."""
import numpy as np
import pytest

from baghera_b200 import synth
from oracle import model_oracle as mo
from oracle.c_oracle import COracle
from helpers import grad_rel_err


def _inputs(d):
    return d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64)


@pytest.mark.parametrize("fix", [False, True])
def test_four_forms_agree(cfg1, fix):
    chi2, ld, cat = _inputs(cfg1)
    ss = mo.SuffStats(chi2, ld, cat, cfg1.G)
    co = COracle(chi2, ld, cat, cfg1.G, cfg1.c, fix)
    Q = synth.jittered_start(cfg1.G + 3, 3, seed=11)
    for q in Q:
        lp_a, g_a = mo.logp_grad_autograd(q, chi2, ld, cat, cfg1.c, fix)
        for lp, g in (mo.logp_grad_closed(q, chi2, ld, cat, cfg1.c, fix),
                      mo.logp_grad_suffstat(q, ss, cfg1.c, fix),
                      co.logp_grad(q)):
            assert abs(lp - lp_a) <= 1e-12 * abs(lp_a)
            assert grad_rel_err(g, g_a) <= 1e-10


def test_extreme_transforms(cfg1):
    """x_W, x_m = +-20, beta < 0, beta huge: closed form == autograd."""
    chi2, ld, cat = _inputs(cfg1)
    rng = np.random.default_rng(5)
    for xW, xm in [(-20.0, 20.0), (20.0, -20.0), (-3.0, 0.0), (5.0, 7.0)]:
        q = np.concatenate([[xW, 0.3, xm], rng.normal(0.0, 2.0, cfg1.G)])
        lp_a, g_a = mo.logp_grad_autograd(q, chi2, ld, cat, cfg1.c)
        lp_b, g_b = mo.logp_grad_closed(q, chi2, ld, cat, cfg1.c)
        assert np.isfinite(lp_a) and abs(lp_b - lp_a) <= 1e-11 * abs(lp_a)
        assert grad_rel_err(g_b, g_a) <= 1e-8


def test_finite_differences(cfg1):
    chi2, ld, cat = _inputs(cfg1)
    q = synth.jittered_start(cfg1.G + 3, 1, seed=2)[0]
    idx = [0, 1, 2, 3, 50, cfg1.G + 2]
    fd = mo.finite_difference(lambda x: mo.logp_grad_closed(x, chi2, ld, cat, cfg1.c)[0], q, idx, h=1e-6)
    _, g = mo.logp_grad_closed(q, chi2, ld, cat, cfg1.c)
    assert np.allclose(fd, g[idx], rtol=2e-5, atol=1e-4)


def test_permutation_invariance(cfg1):
    """Q6: the reference gathers through an unsorted cat; logp/grad do not depend on SNP order."""
    chi2, ld, cat = _inputs(cfg1)
    q = synth.jittered_start(cfg1.G + 3, 1, seed=4)[0]
    order = np.argsort(cat, kind="stable")
    lp1, g1 = mo.logp_grad_closed(q, chi2, ld, cat, cfg1.c)
    lp2, g2 = mo.logp_grad_closed(q, chi2[order], ld[order], cat[order], cfg1.c)
    assert abs(lp1 - lp2) <= 1e-12 * abs(lp1) and grad_rel_err(g2, g1) <= 1e-10


def test_gw_forms_agree():
    d = synth.make_arrays(30_000, 1, seed=7)
    chi2, ld = d.chi2.astype(np.float64), d.ld.astype(np.float64)
    co = COracle(chi2, ld, d.gene_id, 1, d.c)
    for q in ([1.0, -0.69], [0.2, 3.0], [1.7, -8.0]):
        a = mo.gw_logp_grad_autograd(q, chi2, ld, d.c)
        b = mo.gw_logp_grad_closed(q, chi2, ld, d.c)
        c_ = co.gw_logp_grad(q)
        for lp, g in (b, c_):
            assert abs(lp - a[0]) <= 1e-12 * abs(a[0]) and grad_rel_err(g, a[1]) <= 1e-10


def test_c_oracle_batch_threads(cfg1):
    chi2, ld, cat = _inputs(cfg1)
    co = COracle(chi2, ld, cat, cfg1.G, cfg1.c)
    Q = synth.jittered_start(cfg1.G + 3, 5, seed=9)
    lp, g = co.logp_grad_batch(Q, n_threads=3)
    for i in range(5):
        l1, g1 = co.logp_grad(Q[i])
        assert lp[i] == l1 and np.array_equal(g[i], g1)


def test_synth_shapes():
    d = synth.make_arrays(20_000, 200)
    assert d.M == 20_000 and d.G == 200 and d.chi2.dtype == np.float32 and d.ld.min() >= 1.0
    counts = np.bincount(d.gene_id, minlength=d.G)
    assert counts.min() >= 1 and counts[-1] == 9000 and d.gene_names[-1] == "NonCoding"
    assert abs(d.c - 361194 / 1290028) < 1e-15


def test_gamma_model_forms_agree():
    """Gamma model (ref: gene_regression.py:229-248): PyMC3-shaped autograd == closed form; finite differences."""
    rng = np.random.default_rng(5)
    M, G = 3000, 12
    cat = rng.integers(0, G, M)
    cat[:G] = np.arange(G)
    ld = 1 + np.round(rng.lognormal(3.5, 1, M), 3)
    n, rate = 361194.0, 1290028.0
    beta = rng.gamma(0.3, 1 / rate, G) + 1e-9
    chi2 = 1.02 + n * ld * beta[cat] + np.sqrt(ld) * rng.normal(size=M)
    for fix in (False, True):
        for trial in range(3):
            q = np.concatenate([[1.0 + 1e-3 * rng.normal(), rng.normal()], np.log(beta) + rng.normal(0, 0.3, G)])
            lp_a, g_a = mo.gamma_logp_grad_autograd(q, chi2, ld, cat, n, rate, fix)
            lp_b, g_b = mo.gamma_logp_grad_closed(q, chi2, ld, cat, n, rate, fix)
            assert abs(lp_a - lp_b) <= 1e-12 * abs(lp_a)
            assert np.max(np.abs(g_a - g_b) / np.maximum(np.abs(g_a), 1e-9 * np.abs(g_a).max())) <= 1e-10
    f = lambda v: mo.gamma_logp_grad_closed(v, chi2, ld, cat, n, rate)[0]
    g_b = mo.gamma_logp_grad_closed(q, chi2, ld, cat, n, rate)[1]
    for idx in (0, 1, 2, G + 1):
        h = 1e-6 if idx != 0 else 1e-7
        fd = (f(q + h * np.eye(G + 2)[idx]) - f(q - h * np.eye(G + 2)[idx])) / (2 * h)
        assert abs(fd - g_b[idx]) <= 2e-5 * max(abs(g_b[idx]), 1.0), (idx, fd, g_b[idx])


def test_density_against_scipy_distributions(cfg1):
    """Independent third-party pin of the density formulas: the prior / likelihood terms evaluated with
    scipy.stats' own implementations of the distributions the reference names (InverseGamma(1,1), Normal(1,1),
    Beta(1,1), Normal(mi, W), Normal(mu_j, sqrt(l_j)); ref: gene_regression.py:78-98) plus the change-of-variables
    terms of the log / logodds transforms must reproduce the oracle's logp (PyMC3 itself cannot run here)."""
    from scipy import stats
    from scipy.special import expit
    chi2, ld, cat = _inputs(cfg1)
    for q in synth.jittered_start(cfg1.G + 3, 3, seed=21):
        xW, e, xm, beta = q[0], q[1], q[2], q[3:]
        W, mi = np.exp(xW), expit(xm)
        lp = stats.invgamma.logpdf(W, a=1.0, scale=1.0) + xW                       # log transform: |dW/dx| = W
        lp += stats.norm.logpdf(e, loc=1.0, scale=1.0)
        lp += stats.beta.logpdf(mi, 1.0, 1.0) + np.log(mi) + np.log1p(-mi)          # logodds: |dmi/dx| = mi (1 - mi)
        lp += stats.norm.logpdf(beta, loc=mi, scale=W).sum()
        lp += stats.norm.logpdf(chi2, loc=cfg1.c * beta[cat] * ld + e, scale=np.sqrt(ld)).sum()
        lp_b, _ = mo.logp_grad_closed(q, chi2, ld, cat, cfg1.c)
        assert abs(lp - lp_b) <= 1e-12 * abs(lp_b)
    # genome-wide model (ref: heritability.py:44-55): e ~ Normal(1,1), mi ~ Beta(1,2)
    for q in synth.jittered_start(2, 3, seed=22, model="gw"):
        e, xm = q
        mi = expit(xm)
        lp = stats.norm.logpdf(e, 1.0, 1.0) + stats.beta.logpdf(mi, 1.0, 2.0) + np.log(mi) + np.log1p(-mi)
        lp += stats.norm.logpdf(chi2, loc=cfg1.c * mi * ld + e, scale=np.sqrt(ld)).sum()
        lp_b, _ = mo.gw_logp_grad_closed(q, chi2, ld, cfg1.c)
        assert abs(lp - lp_b) <= 1e-12 * abs(lp_b)
    # gamma model (ref: gene_regression.py:229-247): e ~ Normal(1, 0.001), beta ~ Gamma(alpha=mi, beta=N_1kG) on exp(x)
    n_pat, rate = 361194.0, 1290028.0
    for q in synth.jittered_start(cfg1.G + 2, 2, seed=23, model="gamma"):
        e, xm, blog = q[0], q[1], q[2:]
        mi, b = expit(xm), np.exp(blog)
        lp = stats.norm.logpdf(e, 1.0, 0.001) + stats.beta.logpdf(mi, 1.0, 1.0) + np.log(mi) + np.log1p(-mi)
        lp += (stats.gamma.logpdf(b, a=mi, scale=1.0 / rate) + blog).sum()
        lp += stats.norm.logpdf(chi2, loc=n_pat * b[cat] * ld + e, scale=np.sqrt(ld)).sum()
        lp_b, _ = mo.gamma_logp_grad_closed(q, chi2, ld, cat, n_pat, rate)
        assert abs(lp - lp_b) <= 1e-11 * abs(lp_b)


def test_density_and_gradient_against_torch_distributions(cfg1):
    """Pin F: torch.distributions builds the SAME transformed model the reference declares to PyMC3 — the priors as
    TransformedDistribution objects on the unconstrained scale (log for the InverseGamma, logit for the Beta: the library
    supplies the log-abs-det-Jacobian, nothing is hand-derived here) — and autograd differentiates it.  logp AND every
    gradient component of the closed-form oracle must agree (ref: gene_regression.py:78-98; pymc3 3.6 transforms
    `log` and `logodds`)."""
    import torch
    from torch import distributions as D
    from torch.distributions import transforms as T
    chi2, ld, cat = _inputs(cfg1)
    s_t, l_t, cat_t = torch.tensor(chi2), torch.tensor(ld), torch.tensor(cat)
    one = torch.tensor(1.0, dtype=torch.float64)
    prior_xW = D.TransformedDistribution(D.InverseGamma(one, one), [T.ExpTransform().inv])           # x_W = log W
    prior_xm = D.TransformedDistribution(D.Beta(one, one), [T.SigmoidTransform().inv])               # x_m = logit mi
    for q in synth.jittered_start(cfg1.G + 3, 3, seed=31):
        x = torch.tensor(q, dtype=torch.float64, requires_grad=True)
        xW, e, xm, beta = x[0], x[1], x[2], x[3:]
        W, mi = torch.exp(xW), torch.sigmoid(xm)
        lp = prior_xW.log_prob(xW) + D.Normal(one, one).log_prob(e) + prior_xm.log_prob(xm)
        lp = lp + D.Normal(mi, W).log_prob(beta).sum()
        lp = lp + D.Normal(cfg1.c * beta[cat_t] * l_t + e, torch.sqrt(l_t)).log_prob(s_t).sum()
        lp.backward()
        lp_b, g_b = mo.logp_grad_closed(q, chi2, ld, cat, cfg1.c)
        assert abs(lp.item() - lp_b) <= 1e-12 * abs(lp_b)
        g = x.grad.numpy()
        scale = np.maximum(np.abs(g_b), 1e-9 * np.abs(g_b).max())
        assert np.max(np.abs(g - g_b) / scale) <= 1e-9
    # genome-wide model: mi ~ Beta(1, 2) on the logit scale (ref: heritability.py:44-55)
    two = torch.tensor(2.0, dtype=torch.float64)
    prior_gw = D.TransformedDistribution(D.Beta(one, two), [T.SigmoidTransform().inv])
    for q in synth.jittered_start(2, 3, seed=32, model="gw"):
        x = torch.tensor(q, dtype=torch.float64, requires_grad=True)
        e, xm = x[0], x[1]
        lp = D.Normal(one, one).log_prob(e) + prior_gw.log_prob(xm)
        lp = lp + D.Normal(cfg1.c * torch.sigmoid(xm) * l_t + e, torch.sqrt(l_t)).log_prob(s_t).sum()
        lp.backward()
        lp_b, g_b = mo.gw_logp_grad_closed(q, chi2, ld, cfg1.c)
        assert abs(lp.item() - lp_b) <= 1e-12 * abs(lp_b)
        assert np.max(np.abs(x.grad.numpy() - g_b) / np.maximum(np.abs(g_b), 1e-9)) <= 1e-9
