"""GPU: the batched NUTS driver (C ABI + CUDA kernels) against the CPU oracle NUTS.

Sampler parity is statistical (the reference never seeds pm.sample, quirk Q9): posterior means of
every coordinate must agree with the oracle's within Monte-Carlo standard error, the adapted
acceptance rate must sit at target_accept = 0.90, and R-hat must be ~1.
"""
import os

import numpy as np
import pytest

from baghera_b200 import stats as st
from baghera_b200 import synth

pytestmark = pytest.mark.gpu


def _setup(M, G, C, seed, data_seed):
    from baghera_b200.engine import Engine
    from baghera_b200.sampler import NutsSampler
    d = synth.make_arrays(M, G, seed=data_seed)
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, C)
    s = NutsSampler(eng, seed=seed)
    return d, eng, s


def test_posterior_matches_oracle_nuts():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "nuts_small.npz"))
    M, G = int(z["M"]), int(z["G"])
    C = 8
    d, eng, s = _setup(M, G, C, seed=7, data_seed=int(z["seed"]))
    s.init(synth.jittered_start(G + 3, C, seed=21))
    out = s.sample(draws=1500, tune=1000)
    tr = out["trace"].cpu().numpy()                      # [draws, D, C]
    tr = np.transpose(tr, (0, 2, 1))                     # [draws, C, D]
    stats = out["stats"]
    post = stats[1000:]
    assert np.all(np.isfinite(tr))
    acc = post[:, 1].mean()
    assert 0.85 < acc < 0.95, acc                        # target_accept = 0.90
    assert post[:, 3].sum() <= 2                         # (almost) no divergences
    assert abs(post[:, 0].mean() - float(z["depth_mean"])) < 0.6
    D = G + 3
    mean, sd = tr.mean(axis=(0, 1)), tr.std(axis=(0, 1))
    ess = np.array([st.effective_n(tr[:, :, k]) for k in range(D)])
    rhat = np.array([st.gelman_rubin(tr[:, :, k]) for k in range(D)])
    assert rhat.max() < 1.02, rhat.max()
    assert ess.min() > 400, ess.min()
    # difference of two independent MC estimates: sd * sqrt(1/ess_gpu + 1/ess_oracle)
    se = np.sqrt(sd ** 2 / ess + z["sd"] ** 2 / z["ess"])
    zscore = (mean - z["mean"]) / se
    assert np.max(np.abs(zscore)) < 4.5, zscore
    assert np.sqrt(np.mean(zscore ** 2)) < 1.6
    assert np.allclose(sd, z["sd"], rtol=0.12)
    eng.close()


def test_sampler_is_reproducible_and_seed_sensitive():
    d, eng, s = _setup(3000, 10, 4, seed=5, data_seed=3)
    q0 = synth.jittered_start(13, 4, seed=1)
    s.init(q0)
    a = s.sample_lockstep(draws=20, tune=30)["trace"].cpu().numpy()
    from baghera_b200.sampler import NutsSampler
    s2 = NutsSampler(eng, seed=5)
    s2.init(q0)
    b = s2.sample_lockstep(draws=20, tune=30)["trace"].cpu().numpy()
    assert np.array_equal(a, b)                          # counter-based RNG + fixed summation order
    s3 = NutsSampler(eng, seed=6)
    s3.init(q0)
    c = s3.sample_lockstep(draws=20, tune=30)["trace"].cpu().numpy()
    assert not np.array_equal(a, c)
    eng.close()


def test_energy_is_conserved_by_the_integrator():
    """With a tiny step size every leaf is accepted with p ~ 1 and trees grow to the depth cap."""
    d, eng, s = _setup(3000, 10, 4, seed=2, data_seed=3)
    s.init(synth.jittered_start(13, 4, seed=1))
    s.sc[0].fill_(1e-4)                                   # step size
    n = s.transition(False, max_treedepth=5)
    eng.torch.cuda.synchronize()
    stt = s.stats.cpu().numpy()[:, :4]
    assert n == 31 and np.all(stt[0] == 5)                # 2^5 - 1 leapfrogs, depth 5 (no U-turn yet)
    assert np.all(stt[1] > 0.999) and np.all(np.abs(stt[7]) < 1e-2)
    eng.close()


def test_bad_start_raises():
    d, eng, s = _setup(3000, 10, 4, seed=2, data_seed=3)
    q0 = synth.jittered_start(13, 4, seed=1)
    q0[2, 5] = np.nan                                     # non-finite start -> PyMC3 raises "Bad initial energy"
    with pytest.raises(ValueError, match="Bad initial energy"):
        s.init(q0)
    eng.close()


def test_full_size_transitions():
    """cfg3 shape: a few tuned transitions at 1.1M x 15k x 64 chains stay finite and adapt."""
    d, eng, s = _setup(1_100_000, 15_000, 64, seed=9, data_seed=20211)
    s.init(synth.jittered_start(eng.D, 64, seed=4))
    lp0 = s.logp.cpu().numpy().copy()
    for i in range(12):
        s.transition(True, 8)
    eng.torch.cuda.synchronize()
    lp = s.logp.cpu().numpy()
    assert np.all(np.isfinite(lp)) and np.all(lp > lp0)   # chains climb from the jittered start
    assert np.all(np.isfinite(s.q.cpu().numpy()))
    eng.close()


@pytest.mark.parametrize("engine", ["1", "2"])
def test_async_machine_reproduces_lockstep_chains(engine, monkeypatch):
    """The persistent per-chain state machines (BGH_TICK=1: nuts_tick_kernel, 2: the fused kernel of bgh_tick2.cu) make the
    same decisions as the lock-step loop: with the same seed all produce the same chains (Philox counters are keyed by each
    chain's own draw index).  The fused kernel sums the per-chain quantities in another order than the lock-step kernels, so
    its positions agree to rounding (1e-11 over a short horizon) and drift apart slowly over many leapfrogs of a chaotic
    integrator; tree depths and tree sizes of all 65 transitions must be EQUAL."""
    from baghera_b200.sampler import NutsSampler
    monkeypatch.setenv("BGH_TICK", engine)
    d, eng, s = _setup(3000, 10, 4, seed=5, data_seed=3)
    q0 = synth.jittered_start(13, 4, seed=1)
    s.init(q0)
    a = s.sample_lockstep(draws=25, tune=40)
    s2 = NutsSampler(eng, seed=5)
    s2.init(q0)
    b = s2.sample(draws=25, tune=40, poll_ticks=16)
    ta, tb = a["trace"].cpu().numpy(), b["trace"].cpu().numpy()
    assert np.array_equal(a["stats"][:, 0], b["stats"][:, 0])          # tree depths
    assert np.array_equal(a["stats"][:, 2], b["stats"][:, 2])          # tree sizes
    assert np.array_equal(a["stats"][:, 3], b["stats"][:, 3])          # divergences
    # (measured with tools/tick2_vs_lockstep.py: the energy difference is 0 for the first transitions, one ulp after the first
    # deep tree, and saturates near 1e-6 relative once the adapted step sizes differ by 1e-6: no jump at any event)
    assert np.allclose(a["stats"][:, 5], b["stats"][:, 5], rtol=1e-12 if engine == "1" else 1e-5)     # step sizes
    if engine == "1":
        assert np.allclose(ta, tb, rtol=1e-12, atol=0)
    else:
        assert np.allclose(ta, tb, rtol=1e-3, atol=1e-5)
    # short horizon, no adaptation: rounding-level agreement
    s3 = NutsSampler(eng, seed=9)
    s3.init(q0)
    c = s3.sample_lockstep(draws=8, tune=0)
    s4 = NutsSampler(eng, seed=9)
    s4.init(q0)
    e = s4.sample(draws=8, tune=0, poll_ticks=16)
    assert np.array_equal(c["stats"][:, :3], e["stats"][:, :3])
    assert np.allclose(c["trace"].cpu().numpy(), e["trace"].cpu().numpy(), rtol=1e-11, atol=0)
    # the async run needs far fewer batched evaluations than the lock-step one
    assert b["ticks"] <= a["leapfrogs"].sum() + 65 * 16
    eng.close()


@pytest.mark.parametrize("C,M,G", [(1, 3000, 10), (3, 3000, 10), (13, 5000, 40), (32, 5000, 40), (100, 20000, 300)])
def test_tick_kernel_chain_geometries(C, M, G, monkeypatch):
    """Every chain-lane geometry of the persistent tick kernel (padded chain counts, several chain blocks,
    rows not a multiple of anything) makes the lock-step kernels' decisions."""
    from baghera_b200.sampler import NutsSampler
    monkeypatch.setenv("BGH_TICK", "2")
    d, eng, s = _setup(M, G, C, seed=11, data_seed=5)
    q0 = synth.jittered_start(G + 3, C, seed=2)
    s.init(q0)
    a = s.sample_lockstep(draws=6, tune=14)
    s2 = NutsSampler(eng, seed=11)
    s2.init(q0)
    b = s2.sample(draws=6, tune=14, poll_ticks=32)
    assert np.array_equal(a["stats"][:, 0], b["stats"][:, 0])          # tree depths
    assert np.array_equal(a["stats"][:, 2], b["stats"][:, 2])          # tree sizes
    # reduction trees differ between the two drivers when chains span several 64-chain blocks: same decisions,
    # positions equal up to accumulated rounding (measured 8e-7 relative on near-zero coordinates at C=100)
    assert np.allclose(a["trace"].cpu().numpy(), b["trace"].cpu().numpy(), rtol=1e-5, atol=1e-7)
    # state handed back in the caller's layout: a lock-step transition can continue from it
    n = s2.transition(False, 6)
    assert n >= 1 and np.all(np.isfinite(s2.q.cpu().numpy()))
    eng.close()


def test_tick_kernel_gw_and_fixed_intercept():
    """The tick kernel is model agnostic: genome-wide model (D = 2) and fix_intercept."""
    from baghera_b200.engine import Engine
    from baghera_b200.sampler import NutsSampler
    d = synth.make_arrays(20000, 1, seed=4)
    eng = Engine(d.chi2, d.ld, None, 1, d.c, 6, model="gw")
    s = NutsSampler(eng, seed=3)
    s.init(synth.jittered_start(2, 6, seed=1, model="gw"))
    out = s.sample(draws=200, tune=200)
    tr = out["trace"].cpu().numpy()
    assert np.all(np.isfinite(tr)) and abs(tr[:, 0, :].mean() - d.truth["e"]) < 0.1
    eng.close()
    d = synth.make_arrays(5000, 20, seed=4)
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, 4, fix_intercept=True)
    s = NutsSampler(eng, seed=3)
    s.init(synth.jittered_start(d.G + 3, 4, seed=1))
    out = s.sample(draws=50, tune=100)
    assert np.all(np.isfinite(out["trace"].cpu().numpy()))
    eng.close()
