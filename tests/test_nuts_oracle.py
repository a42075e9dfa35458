"""CPU: the oracle NUTS (PyMC3 3.6 restatement) recovers known posteriors; stats helpers behave."""
import numpy as np

from baghera_b200 import stats as st
from oracle import nuts_oracle as no


def test_oracle_nuts_known_gaussian():
    """Independent Gaussian target with very different scales: means / sds recovered within MC error."""
    mu = np.array([0.0, 3.0, -2.0, 10.0, 0.5])
    sd = np.array([1.0, 0.1, 5.0, 0.01, 2.0])

    def f(q):
        z = (q - mu) / sd
        return -0.5 * float(z @ z), -z / sd

    rng = np.random.default_rng(0)
    starts, _ = no.jitter_start(mu, 4, rng)
    trace, stats, _ = no.sample(f, starts, draws=1500, tune=1000, seed=3)
    m, s = trace.mean(axis=(0, 1)), trace.std(axis=(0, 1))
    ess = np.array([st.effective_n(trace[:, :, k]) for k in range(5)])
    assert np.all(np.abs(m - mu) < 5.0 * sd / np.sqrt(ess))
    assert np.allclose(s, sd, rtol=0.1)
    acc = np.mean([x["mean_tree_accept"] for c in stats for x in c[1000:]])
    assert 0.8 < acc < 0.98          # target_accept = 0.90
    assert max(st.gelman_rubin(trace[:, :, k]) for k in range(5)) < 1.02


def test_stats_helpers():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(2000, 4))
    assert abs(st.gelman_rubin(x) - 1.0) < 0.01
    assert 6000 < st.effective_n(x) < 10000
    lo, hi = st.hpd(x, 0.05)
    assert -2.2 < lo < -1.7 and 1.7 < hi < 2.2
    assert abs(st.mc_error(x) - 1.0 / np.sqrt(8000)) < 0.005
    # strongly autocorrelated AR(1) chains have far fewer effective draws
    y = np.zeros((2000, 4))
    for i in range(1, 2000):
        y[i] = 0.95 * y[i - 1] + rng.normal(size=4)
    assert st.effective_n(y) < 600
    df = st.summary({"a": x, "b": y})
    assert list(df.columns) == ["mean", "sd", "mc_error", "hpd_2.5", "hpd_97.5", "median", "5", "50", "95", "n_eff", "Rhat"]
    assert list(df.index) == ["a", "b"]
