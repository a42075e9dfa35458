"""CPU restatement of PyMC3 3.6's NUTS sampler as the reference calls it.

TEST INFRASTRUCTURE / REPORTED CPU BASELINE ONLY (see oracle/model_oracle.py header).

PARITY UNPINNED: ``pm.sample(SWEEPS, tune=TUNE, chains=CHAINS, cores=CORES,
nuts_kwargs=dict(target_accept=0.90))`` (ref: baghera_tool/gene_regression.py:100-106) runs inside
pymc3==3.6, which is not installable here.  This module restates, from the published 3.6 sources
[recalled; SURVEY §8a], the pieces that call produces with default arguments:

  sampling.init_nuts('jitter+adapt_diag')            -> :func:`jitter_start`, :class:`QuadPotentialDiagAdapt`
  step_methods/hmc/quadpotential.py                  -> :class:`QuadPotentialDiagAdapt`, :class:`_WeightedVariance`
  step_methods/hmc/step_size.py                      -> :class:`DualAverageAdaptation`
  step_methods/hmc/integration.py  CpuLeapfrogIntegrator -> :func:`leapfrog`
  step_methods/hmc/nuts.py  NUTS / _Tree (recursive multinomial NUTS, biased progressive top level)
  step_methods/hmc/base_hmc.py  BaseHMC.astep        -> :meth:`NutsChain.step`

The reference passes no random_seed (quirk Q9), so sampler-level parity is statistical only.
"""
from __future__ import annotations

from collections import namedtuple

import numpy as np

State = namedtuple("State", "q p v q_grad energy model_logp")
Proposal = namedtuple("Proposal", "q q_grad energy p_accept logp")
Subtree = namedtuple("Subtree", "left right p_sum proposal log_size accept_sum n_proposals")


class _WeightedVariance:
    def __init__(self, nelem, initial_mean=None, initial_variance=None, initial_weight=0):
        self.w_sum = float(initial_weight)
        self.w_sum2 = float(initial_weight) ** 2
        self.mean = np.zeros(nelem) if initial_mean is None else np.array(initial_mean, dtype="d", copy=True)
        self.raw_var = np.zeros(nelem) if initial_variance is None else np.array(initial_variance, dtype="d", copy=True)
        self.raw_var[:] *= self.w_sum

    def add_sample(self, x, weight):
        x = np.asarray(x)
        self.w_sum += weight
        self.w_sum2 += weight * weight
        prop = weight / self.w_sum
        old_diff = x - self.mean
        self.mean[:] += prop * old_diff
        new_diff = x - self.mean
        self.raw_var[:] += weight * old_diff * new_diff

    def current_variance(self):
        if self.w_sum < 1:
            raise ValueError("Can not compute variance without samples.")
        return self.raw_var / self.w_sum


class QuadPotentialDiagAdapt:
    def __init__(self, n, initial_mean, initial_diag=None, initial_weight=0, adaptation_window=101, rng=None):
        if initial_diag is None:
            initial_diag = np.ones(n)
            initial_weight = 1
        self._n = n
        self._var = np.array(initial_diag, dtype="d", copy=True)
        self._stds = np.sqrt(initial_diag)
        self._inv_stds = 1.0 / self._stds
        self._foreground_var = _WeightedVariance(n, initial_mean, initial_diag, initial_weight)
        self._background_var = _WeightedVariance(n)
        self._n_samples = 0
        self.adaptation_window = adaptation_window
        self.rng = rng or np.random.default_rng()

    def velocity(self, x):
        return self._var * x

    def energy(self, x, velocity=None):
        if velocity is not None:
            return 0.5 * x.dot(velocity)
        return 0.5 * x.dot(self._var * x)

    def random(self):
        return self._inv_stds * self.rng.standard_normal(self._n)

    def update(self, sample, grad, tune):
        if not tune:
            return
        window = self.adaptation_window
        self._foreground_var.add_sample(sample, weight=1)
        self._background_var.add_sample(sample, weight=1)
        self._var = self._foreground_var.current_variance()
        self._stds = np.sqrt(self._var)
        self._inv_stds = 1.0 / self._stds
        if self._n_samples > 0 and self._n_samples % window == 0:
            self._foreground_var = self._background_var
            self._background_var = _WeightedVariance(self._n)
        self._n_samples += 1


class DualAverageAdaptation:
    def __init__(self, initial_step, target, gamma, k, t0):
        self._log_step = np.log(initial_step)
        self._log_bar = self._log_step
        self._target = target
        self._hbar = 0.0
        self._k = k
        self._t0 = t0
        self._count = 1
        self._mu = np.log(10 * initial_step)
        self._gamma = gamma

    def current(self, tune):
        return np.exp(self._log_step) if tune else np.exp(self._log_bar)

    def update(self, accept_stat, tune):
        if not tune:
            return
        count, k, t0 = self._count, self._k, self._t0
        w = 1.0 / (count + t0)
        self._hbar = (1 - w) * self._hbar + w * (self._target - accept_stat)
        self._log_step = self._mu - self._hbar * np.sqrt(count) / self._gamma
        mk = count ** -k
        self._log_bar = mk * self._log_step + (1 - mk) * self._log_bar
        self._count += 1


def _logbern(log_p, rng):
    if np.isnan(log_p):
        raise FloatingPointError("log_p can't be nan.")
    return np.log(rng.uniform()) < log_p


class _Tree:
    def __init__(self, ndim, chain, start, step_size, emax):
        self.chain = chain
        self.start = start
        self.step_size = step_size
        self.emax = emax
        self.start_energy = start.energy
        self.left = self.right = start
        self.proposal = Proposal(start.q, start.q_grad, start.energy, 1.0, start.model_logp)
        self.depth = 0
        self.log_size = 0.0
        self.accept_sum = 0.0
        self.n_proposals = 0
        self.p_sum = start.p.copy()
        self.max_energy_change = 0.0

    def extend(self, direction):
        rng = self.chain.rng
        if direction > 0:
            tree, diverging, turning = self._build_subtree(self.right, self.depth, self.step_size)
            self.right = tree.right
        else:
            tree, diverging, turning = self._build_subtree(self.left, self.depth, -self.step_size)
            self.left = tree.right
        self.depth += 1
        self.accept_sum += tree.accept_sum
        self.n_proposals += tree.n_proposals
        if diverging or turning:
            return diverging, turning
        size1, size2 = self.log_size, tree.log_size
        if _logbern(size2 - size1, rng):
            self.proposal = tree.proposal
        self.log_size = np.logaddexp(self.log_size, tree.log_size)
        self.p_sum[:] += tree.p_sum
        left, right = self.left, self.right
        p_sum = self.p_sum
        turning = (p_sum.dot(left.v) <= 0) or (p_sum.dot(right.v) <= 0)
        return diverging, turning

    def _single_step(self, left, epsilon):
        right = self.chain.leapfrog(epsilon, left)
        energy_change = right.energy - self.start_energy
        if np.isnan(energy_change):
            energy_change = np.inf
        if np.abs(energy_change) > np.abs(self.max_energy_change):
            self.max_energy_change = energy_change
        if np.abs(energy_change) < self.emax:
            p_accept = min(1, np.exp(-energy_change))
            log_size = -energy_change
            proposal = Proposal(right.q, right.q_grad, right.energy, p_accept, right.model_logp)
            return Subtree(right, right, right.p, proposal, log_size, p_accept, 1), False, False
        return Subtree(None, None, None, None, -np.inf, 0, 1), True, False

    def _build_subtree(self, left, depth, epsilon):
        if depth == 0:
            return self._single_step(left, epsilon)
        tree1, diverging, turning = self._build_subtree(left, depth - 1, epsilon)
        if diverging or turning:
            return tree1, diverging, turning
        tree2, diverging, turning = self._build_subtree(tree1.right, depth - 1, epsilon)
        left, right = tree1.left, tree2.right
        if not (diverging or turning):
            p_sum = tree1.p_sum + tree2.p_sum
            turning = (p_sum.dot(left.v) <= 0) or (p_sum.dot(right.v) <= 0)
            log_size = np.logaddexp(tree1.log_size, tree2.log_size)
            if _logbern(tree2.log_size - log_size, self.chain.rng):
                proposal = tree2.proposal
            else:
                proposal = tree1.proposal
        else:
            p_sum = tree1.p_sum
            log_size = tree1.log_size
            proposal = tree1.proposal
        accept_sum = tree1.accept_sum + tree2.accept_sum
        n_proposals = tree1.n_proposals + tree2.n_proposals
        return Subtree(left, right, p_sum, proposal, log_size, accept_sum, n_proposals), diverging, turning


class NutsChain:
    """One chain of pm.sample's NUTS step method (BaseHMC defaults + target_accept)."""

    def __init__(self, logp_dlogp, start, initial_mean, rng, target_accept=0.90, max_treedepth=10,
                 early_max_treedepth=8, step_scale=0.25, gamma=0.05, k=0.75, t0=10, emax=1000.0):
        self.f = logp_dlogp
        self.rng = rng
        self.q = np.array(start, dtype="d", copy=True)
        n = self.q.shape[0]
        self.potential = QuadPotentialDiagAdapt(n, initial_mean, np.ones(n), 10, rng=rng)
        self.step_adapt = DualAverageAdaptation(step_scale / n ** 0.25, target_accept, gamma, k, t0)
        self.max_treedepth = max_treedepth
        self.early_max_treedepth = early_max_treedepth
        self.emax = emax
        self.tune = True
        self.iter_count = 0
        self.n_evals = 0

    def compute_state(self, q, p):
        logp, dlogp = self.f(q)
        self.n_evals += 1
        v = self.potential.velocity(p)
        kinetic = self.potential.energy(p, velocity=v)
        return State(q, p, v, dlogp, kinetic - logp, logp)

    def leapfrog(self, epsilon, state):
        q, p, v, q_grad, energy, logp = state
        dt = 0.5 * epsilon
        p_new = p + dt * q_grad
        v_new = self.potential.velocity(p_new)
        q_new = q + epsilon * v_new
        logp, q_new_grad = self.f(q_new)
        self.n_evals += 1
        p_new = p_new + dt * q_new_grad
        v_new = self.potential.velocity(p_new)
        kinetic = 0.5 * p_new.dot(v_new)
        return State(q_new, p_new, v_new, q_new_grad, kinetic - logp, logp)

    def step(self):
        p0 = self.potential.random()
        start = self.compute_state(self.q, p0)
        if not np.isfinite(start.energy):
            raise ValueError("Bad initial energy: %s. The model might be misspecified." % start.energy)
        adapt_step = self.tune
        step_size = self.step_adapt.current(adapt_step)
        max_treedepth = self.early_max_treedepth if (self.tune and self.iter_count < 200) else self.max_treedepth
        tree = _Tree(len(p0), self, start, step_size, self.emax)
        diverging = False
        for _ in range(max_treedepth):
            direction = _logbern(np.log(0.5), self.rng) * 2 - 1
            diverging, turning = tree.extend(direction)
            if diverging or turning:
                break
        accept_stat = tree.accept_sum / tree.n_proposals
        self.step_adapt.update(accept_stat, adapt_step)
        self.potential.update(tree.proposal.q, tree.proposal.q_grad, self.tune)
        self.q = tree.proposal.q
        self.iter_count += 1
        return dict(depth=tree.depth, mean_tree_accept=accept_stat, tree_size=tree.n_proposals, diverging=bool(diverging),
                    energy=tree.proposal.energy, step_size=step_size, model_logp=tree.proposal.logp)

    def stop_tuning(self):
        self.tune = False


def jitter_start(test_point, chains, rng):
    """init='jitter+adapt_diag': test point + U(-1, 1) per coordinate and chain; returns (starts, mean)."""
    tp = np.asarray(test_point, dtype="d")
    starts = tp[None, :] + (2.0 * rng.random((chains, tp.shape[0])) - 1.0)
    return starts, starts.mean(axis=0)


def sample(logp_dlogp, starts, draws, tune, seed=0, target_accept=0.90):
    """pm.sample(draws, tune=tune, chains=len(starts)): returns trace [draws, chains, D] and stats."""
    starts = np.asarray(starts, dtype="d")
    chains, D = starts.shape
    mean = starts.mean(axis=0)
    trace = np.empty((draws, chains, D))
    stats = []
    n_evals = 0
    for c in range(chains):
        rng = np.random.default_rng([seed, c])
        ch = NutsChain(logp_dlogp, starts[c], mean, rng, target_accept=target_accept)
        st = []
        for i in range(tune + draws):
            if i == tune:
                ch.stop_tuning()
            s = ch.step()
            st.append(s)
            if i >= tune:
                trace[i - tune, c] = ch.q
        stats.append(st)
        n_evals += ch.n_evals
    return trace, stats, n_evals
