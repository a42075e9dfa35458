"""CPU oracle for BAGHERA's gene-heritability log-density and gradient.

TEST INFRASTRUCTURE ONLY.  Nothing under ``baghera_b200/`` imports this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs as the checker / reported baseline, never as the product path.

PARITY UNPINNED: the reference's arithmetic for this path lives in the third-party
dependencies ``pymc3==3.6`` (ref: requirements.txt:7, meta.yaml:25,36) and ``theano``,
which are neither under /root/reference nor installable here, and the reference ships no
tests or golden vectors.  This file therefore restates the *published* PyMC3-3.6 formulas
at the reference's own call sites, in three mutually independent forms that the tests
require to agree to <= 1e-12 relative:

  (A) ``logp_pymc3_shaped``  – term-for-term transcription of what PyMC3 builds from
      ref: baghera_tool/gene_regression.py:77-98, differentiated by torch autograd;
  (B) ``logp_grad_closed``   – hand-derived closed-form logp and gradient (numpy);
  (C) ``logp_grad_suffstat`` – exact per-gene sufficient-statistics identity (numpy).

Parameter vector layout (PyMC3 ``ArrayOrdering`` = variable creation order,
ref: baghera_tool/gene_regression.py:78-81):

    gene model : q = [W_log__, e, mi_logodds__, beta_0 .. beta_{G-1}]     D = G + 3
    gw model   : q = [e, mi_logodds__]   (ref: baghera_tool/heritability.py:44-49)  D = 2
"""
from __future__ import annotations

import math

import numpy as np

LOG_2PI = math.log(2.0 * math.pi)


# --------------------------------------------------------------------------------------
# (A) PyMC3-shaped, autograd
# --------------------------------------------------------------------------------------
def logp_pymc3_shaped(q, chi2, ld, cat, c, fix_intercept=False):
    """logp(q) built the way PyMC3 3.6 builds it from ref: gene_regression.py:77-98.

    q: torch float64[D] (requires_grad allowed). chi2/ld: float64[M]; cat: int64[M] (any order).
    Formula sources [PyMC3-3.6, recalled]: distributions/continuous.py, transforms.py.
    """
    import torch

    G = q.shape[0] - 3
    W_log, e, mi_lo, beta = q[0], q[1], q[2], q[3:]
    # transforms.log.backward / transforms.logodds.backward
    W = torch.exp(W_log)
    mi = torch.sigmoid(mi_lo)
    lp = q.new_zeros(())
    # W = pm.InverseGamma("W", alpha=1.0, beta=1.0)                           (:78)
    alpha, b = 1.0, 1.0
    lp = lp + (alpha * math.log(b) - math.lgamma(alpha) - b / W + (-alpha - 1.0) * torch.log(W))
    lp = lp + W_log                                   # log.jacobian_det(x) = x
    # e = pm.Normal("e", mu=1, sd=1)                                          (:79)
    tau = 1.0
    lp = lp + (-tau * (e - 1.0) ** 2 + math.log(tau / math.pi / 2.0)) / 2.0
    # mi = pm.Beta("mi", 1, 1)                                                (:80)
    a_, b_ = 1.0, 1.0
    betaln = math.lgamma(a_) + math.lgamma(b_) - math.lgamma(a_ + b_)
    lp = lp + ((a_ - 1.0) * torch.log(mi) + (b_ - 1.0) * torch.log1p(-mi) - betaln)
    # logodds.jacobian_det(x) = log(sigmoid'(x)) = -2*softplus(-x) - x
    lp = lp + (-2.0 * torch.nn.functional.softplus(-mi_lo) - mi_lo)
    # beta = pm.Normal("beta", mu=mi, sd=W, shape=idx)                        (:81)
    tau_b = W ** -2.0
    lp = lp + torch.sum((-tau_b * (beta - mi) ** 2 + torch.log(tau_b / math.pi / 2.0)) / 2.0)
    # fixed_variable = pm.Normal("fxd", mu=(n/N_1kG)*beta[cat]*l + e, sd=sqrt(l), observed=stats)   (:93-98)
    icpt = 1.0 if fix_intercept else e                                        # (:85-90)
    mu = c * beta[cat] * ld + icpt
    sd = torch.sqrt(ld)
    tau_j = sd ** -2.0
    lp = lp + torch.sum((-tau_j * (chi2 - mu) ** 2 + torch.log(tau_j / math.pi / 2.0)) / 2.0)
    return lp


def logp_grad_autograd(q, chi2, ld, cat, c, fix_intercept=False):
    import torch

    qt = torch.tensor(np.asarray(q, dtype=np.float64), requires_grad=True)
    lp = logp_pymc3_shaped(qt, torch.as_tensor(chi2, dtype=torch.float64),
                           torch.as_tensor(ld, dtype=torch.float64),
                           torch.as_tensor(np.asarray(cat, dtype=np.int64)), float(c), fix_intercept)
    (g,) = torch.autograd.grad(lp, qt)
    return float(lp.detach()), g.numpy()


def gw_logp_pymc3_shaped(q, chi2, ld, c):
    """Genome-wide model, ref: baghera_tool/heritability.py:44-55 (free intercept branch)."""
    import torch

    e, mi_lo = q[0], q[1]
    mi = torch.sigmoid(mi_lo)
    lp = (-(e - 1.0) ** 2 + math.log(1.0 / math.pi / 2.0)) / 2.0               # e ~ Normal(1,1)  (:48)
    a_, b_ = 1.0, 2.0                                                          # mi ~ Beta(1,2)   (:49)
    betaln = math.lgamma(a_) + math.lgamma(b_) - math.lgamma(a_ + b_)
    lp = lp + ((a_ - 1.0) * torch.log(mi) + (b_ - 1.0) * torch.log1p(-mi) - betaln)
    lp = lp + (-2.0 * torch.nn.functional.softplus(-mi_lo) - mi_lo)
    mu = c * mi * ld + e                                                       # (:50-55)
    tau_j = torch.sqrt(ld) ** -2.0
    lp = lp + torch.sum((-tau_j * (chi2 - mu) ** 2 + torch.log(tau_j / math.pi / 2.0)) / 2.0)
    return lp


def gw_logp_grad_autograd(q, chi2, ld, c):
    import torch

    qt = torch.tensor(np.asarray(q, dtype=np.float64), requires_grad=True)
    lp = gw_logp_pymc3_shaped(qt, torch.as_tensor(chi2, dtype=torch.float64),
                              torch.as_tensor(ld, dtype=torch.float64), float(c))
    (g,) = torch.autograd.grad(lp, qt)
    return float(lp.detach()), g.numpy()


# --------------------------------------------------------------------------------------
# (B) closed form (numpy) – follows the Theano op sequence of SURVEY §3.2:
#     gather -> elementwise -> sums -> scatter-add (np.bincount)
# --------------------------------------------------------------------------------------
def lik_const(ld):
    """sum_j -1/2 log(2 pi l_j): the chain-independent normalising constant of (:93-98)."""
    ld = np.asarray(ld, dtype=np.float64)
    return float(-0.5 * np.sum(np.log(ld)) - 0.5 * ld.shape[0] * LOG_2PI)


def logp_grad_closed(q, chi2, ld, cat, c, fix_intercept=False, const=None):
    q = np.asarray(q, dtype=np.float64)
    G = q.shape[0] - 3
    xW, e, xm = q[0], q[1], q[2]
    beta = q[3:]
    W = math.exp(xW)
    mi = 1.0 / (1.0 + math.exp(-xm))
    icpt = 1.0 if fix_intercept else e
    eta = beta[cat]                                    # AdvancedSubtensor1 (gather)
    r = chi2 - icpt - c * ld * eta                     # Elemwise{Composite}
    rw = r / ld
    if const is None:
        const = lik_const(ld)
    ll = -0.5 * float(np.sum(r * rw)) + const          # Sum{acc_dtype=float64}
    d = beta - mi
    S1 = float(np.sum(d))
    S2 = float(np.sum(d * d))
    lp = (-1.0 / W - 2.0 * math.log(W)) + xW
    lp += -0.5 * (e - 1.0) ** 2 - 0.5 * LOG_2PI
    lp += math.log(mi) + math.log1p(-mi)
    lp += -S2 / (2.0 * W * W) - G * math.log(W) - 0.5 * G * LOG_2PI
    lp += ll
    g = np.empty_like(q)
    g[0] = 1.0 / W - 1.0 + S2 / (W * W) - G
    g[1] = -(e - 1.0) + (0.0 if fix_intercept else float(np.sum(rw)))
    g[2] = 1.0 - 2.0 * mi + mi * (1.0 - mi) * S1 / (W * W)
    g[3:] = -d / (W * W) + c * np.bincount(cat, weights=r, minlength=G)   # AdvancedIncSubtensor1
    return lp, g


def gw_logp_grad_closed(q, chi2, ld, c, const=None):
    q = np.asarray(q, dtype=np.float64)
    e, xm = q[0], q[1]
    mi = 1.0 / (1.0 + math.exp(-xm))
    r = chi2 - e - c * ld * mi
    rw = r / ld
    if const is None:
        const = lik_const(ld)
    lp = -0.5 * (e - 1.0) ** 2 - 0.5 * LOG_2PI
    lp += math.log1p(-mi) + math.log(2.0)              # Beta(1,2).logp = log(1-mi) - betaln(1,2) = log(1-mi)+log 2
    lp += math.log(mi) + math.log1p(-mi)               # logodds jacobian
    lp += -0.5 * float(np.sum(r * rw)) + const
    g = np.empty(2)
    g[0] = -(e - 1.0) + float(np.sum(rw))
    dmi = c * float(np.sum(r))                          # d ll / d mi  (l_j cancels)
    g[1] = (dmi - 1.0 / (1.0 - mi)) * mi * (1.0 - mi) + (1.0 - 2.0 * mi)
    return lp, g


# --------------------------------------------------------------------------------------
# (C) sufficient statistics (exact identity)
# --------------------------------------------------------------------------------------
class SuffStats:
    def __init__(self, chi2, ld, cat, G):
        chi2 = np.asarray(chi2, dtype=np.float64)
        ld = np.asarray(ld, dtype=np.float64)
        self.G = G
        self.n = np.bincount(cat, minlength=G).astype(np.float64)
        self.A = np.bincount(cat, weights=ld, minlength=G)
        self.B = np.bincount(cat, weights=chi2, minlength=G)
        self.S1 = float(np.sum(1.0 / ld))
        self.S2 = float(np.sum(chi2 / ld))
        self.S3 = float(np.sum(chi2 * chi2 / ld))
        self.const = lik_const(ld)


def logp_grad_suffstat(q, ss: SuffStats, c, fix_intercept=False):
    q = np.asarray(q, dtype=np.float64)
    G = ss.G
    xW, e, xm = q[0], q[1], q[2]
    beta = q[3:]
    W = math.exp(xW)
    mi = 1.0 / (1.0 + math.exp(-xm))
    ic = 1.0 if fix_intercept else e
    cb = c * beta
    quad = ss.S3 - 2.0 * ic * ss.S2 + ic * ic * ss.S1 + float(
        np.sum(-2.0 * cb * ss.B + 2.0 * ic * cb * ss.n + cb * cb * ss.A))
    d = beta - mi
    S1, S2 = float(np.sum(d)), float(np.sum(d * d))
    lp = (-1.0 / W - 2.0 * math.log(W)) + xW
    lp += -0.5 * (e - 1.0) ** 2 - 0.5 * LOG_2PI
    lp += math.log(mi) + math.log1p(-mi)
    lp += -S2 / (2.0 * W * W) - G * math.log(W) - 0.5 * G * LOG_2PI
    lp += -0.5 * quad + ss.const
    g = np.empty_like(q)
    g[0] = 1.0 / W - 1.0 + S2 / (W * W) - G
    g[1] = -(e - 1.0) + (0.0 if fix_intercept else (ss.S2 - ic * ss.S1 - float(np.sum(cb * ss.n))))
    g[2] = 1.0 - 2.0 * mi + mi * (1.0 - mi) * S1 / (W * W)
    g[3:] = -d / (W * W) + c * (ss.B - ic * ss.n - cb * ss.A)
    return lp, g


# --------------------------------------------------------------------------------------
# helpers shared by tests / bench
# --------------------------------------------------------------------------------------
def batch_closed(Q, chi2, ld, cat, c, fix_intercept=False):
    """Q: float64[C, D] (chain-major). Returns (logp[C], grad[C, D])."""
    Q = np.asarray(Q, dtype=np.float64)
    const = lik_const(ld)
    lps = np.empty(Q.shape[0])
    gs = np.empty_like(Q)
    for i in range(Q.shape[0]):
        lps[i], gs[i] = logp_grad_closed(Q[i], chi2, ld, cat, c, fix_intercept, const)
    return lps, gs


def finite_difference(fun, q, idx, h=1e-5):
    """Central differences of a scalar function at coordinates ``idx``."""
    q = np.asarray(q, dtype=np.float64)
    out = np.empty(len(idx))
    for k, i in enumerate(idx):
        qp, qm = q.copy(), q.copy()
        step = h * max(1.0, abs(q[i]))
        qp[i] += step
        qm[i] -= step
        out[k] = (fun(qp) - fun(qm)) / (2.0 * step)
    return out


GW_FIX_LO, GW_FIX_HI = 0.9999, 1.0000001     # ref: baghera_tool/heritability.py:46


def gw_fix_logp_grad_closed(q, chi2, ld, c):
    """gw model with ``e = pm.Uniform("e", 0.9999, 1.0000001)`` (ref: heritability.py:45-46).

    PyMC3 samples it through the interval transform: e = lo + (hi-lo)*sigmoid(x0);
    Uniform.logp = -log(hi-lo); interval.jacobian_det = log(hi-lo) + log(sigmoid'(x0)).
    """
    q = np.asarray(q, dtype=np.float64)
    x0, xm = q[0], q[1]
    sg = 1.0 / (1.0 + math.exp(-x0))
    e = GW_FIX_LO + (GW_FIX_HI - GW_FIX_LO) * sg
    mi = 1.0 / (1.0 + math.exp(-xm))
    r = chi2 - e - c * ld * mi
    rw = r / ld
    lp = math.log(sg) + math.log1p(-sg)
    lp += math.log1p(-mi) + math.log(2.0)
    lp += math.log(mi) + math.log1p(-mi)
    lp += -0.5 * float(np.sum(r * rw)) + lik_const(ld)
    g = np.empty(2)
    g[0] = float(np.sum(rw)) * (GW_FIX_HI - GW_FIX_LO) * sg * (1.0 - sg) + (1.0 - 2.0 * sg)
    g[1] = (c * float(np.sum(r)) - 1.0 / (1.0 - mi)) * mi * (1.0 - mi) + (1.0 - 2.0 * mi)
    return lp, g


# --------------------------------------------------------------------------------------
# Gamma model (ref: baghera_tool/gene_regression.py:229-248)
#   e ~ Normal(1, sd=0.001); mi ~ Beta(1,1); beta ~ Gamma(alpha=mi, beta=N_1kG, shape=G)
#   stats_j ~ Normal(n_patients * beta[cat_j] * l_j + e, sd=sqrt(l_j))
# free vector (creation order :229-231): q = [e, mi_logodds__, beta_log___0 .. G-1], D = G + 2.
# PyMC3 3.6 [recalled]: Gamma.logp = -gammaln(a) + a log(b) - b x + (a - 1) log(x); log transform jac = x.
# --------------------------------------------------------------------------------------
GAMMA_E_SD = 0.001


def gamma_logp_pymc3_shaped(q, chi2, ld, cat, n_patients, rate, fix_intercept=False):
    """(A) term-for-term, torch (autograd friendly)."""
    import torch

    e, mi_lo, blog = q[0], q[1], q[2:]
    mi = torch.sigmoid(mi_lo)
    beta = torch.exp(blog)
    lp = q.new_zeros(())
    tau = 1.0 / GAMMA_E_SD ** 2
    lp = lp + (-tau * (e - 1.0) ** 2 + math.log(tau / math.pi / 2.0)) / 2.0                 # Normal(1, 0.001)   (:229)
    a_, b_ = 1.0, 1.0
    betaln = math.lgamma(a_) + math.lgamma(b_) - math.lgamma(a_ + b_)
    lp = lp + ((a_ - 1.0) * torch.log(mi) + (b_ - 1.0) * torch.log1p(-mi) - betaln)          # Beta(1,1)          (:230)
    lp = lp + torch.log(mi) + torch.log1p(-mi)                                                # logodds jacobian
    lp = lp + torch.sum(-torch.lgamma(mi) + mi * math.log(rate) - rate * beta + (mi - 1.0) * torch.log(beta))   # (:231)
    lp = lp + torch.sum(blog)                                                                 # log jacobian
    chi2_t = torch.as_tensor(chi2, dtype=q.dtype)
    ld_t = torch.as_tensor(ld, dtype=q.dtype)
    cat_t = torch.as_tensor(cat, dtype=torch.long)
    icpt = q.new_ones(()) if fix_intercept else e
    mu = n_patients * beta[cat_t] * ld_t + icpt                                               # (:234-247)
    tau_j = 1.0 / ld_t
    lp = lp + torch.sum((-tau_j * (chi2_t - mu) ** 2 + torch.log(tau_j / math.pi / 2.0)) / 2.0)
    return lp


def gamma_logp_grad_autograd(q, chi2, ld, cat, n_patients, rate, fix_intercept=False):
    import torch

    qt = torch.tensor(np.asarray(q, dtype=np.float64), dtype=torch.float64, requires_grad=True)
    lp = gamma_logp_pymc3_shaped(qt, chi2, ld, cat, n_patients, rate, fix_intercept)
    (g,) = torch.autograd.grad(lp, qt)
    return float(lp.detach()), g.numpy()


def gamma_logp_grad_closed(q, chi2, ld, cat, n_patients, rate, fix_intercept=False):
    """(B) hand-derived closed form (numpy + scipy.special)."""
    from scipy.special import digamma, gammaln

    q = np.asarray(q, dtype=np.float64)
    G = q.shape[0] - 2
    e, xm, x = q[0], q[1], q[2:]
    mi = 1.0 / (1.0 + np.exp(-xm))
    beta = np.exp(x)
    icpt = 1.0 if fix_intercept else e
    r = chi2 - icpt - n_patients * ld * beta[cat]
    tau = 1.0 / GAMMA_E_SD ** 2
    lp = -0.5 * tau * (e - 1.0) ** 2 + 0.5 * (math.log(tau) - LOG_2PI)
    lp += math.log(mi) + math.log1p(-mi)
    lp += G * (-gammaln(mi) + mi * math.log(rate)) - rate * beta.sum() + (mi - 1.0) * x.sum() + x.sum()
    lp += -0.5 * np.sum(r * r / ld) + lik_const(ld)
    g = np.empty_like(q)
    g[0] = -tau * (e - 1.0) + (0.0 if fix_intercept else np.sum(r / ld))
    g[1] = (1.0 - 2.0 * mi) + mi * (1.0 - mi) * (G * (math.log(rate) - digamma(mi)) + x.sum())
    sums = np.bincount(cat, weights=r, minlength=G)
    g[2:] = beta * (n_patients * sums - rate) + mi
    return float(lp), g


def gamma_batch_closed(Q, chi2, ld, cat, n_patients, rate, fix_intercept=False):
    out = [gamma_logp_grad_closed(q, chi2, ld, cat, n_patients, rate, fix_intercept) for q in Q]
    return np.array([o[0] for o in out]), np.stack([o[1] for o in out])
