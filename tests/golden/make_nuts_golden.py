"""Generates tests/golden/nuts_small.npz: posterior summary of a small gene model from the CPU oracle
NUTS (oracle/nuts_oracle.py + the C logp/grad oracle).  Run once; the .npz is committed.

The GPU sampler test compares its posterior means with these within Monte-Carlo error.  Parity
against PyMC3 itself is unpinned (pymc3==3.6 cannot be installed; it also runs unseeded, Q9).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baghera_b200 import synth                    # noqa: E402
from baghera_b200 import stats as st              # noqa: E402
from oracle import nuts_oracle as no              # noqa: E402
from oracle.c_oracle import COracle               # noqa: E402

M, G, SEED = 4000, 16, 404
CHAINS, TUNE, DRAWS = 4, 1000, 2500
d = synth.make_arrays(M, G, seed=SEED)
co = COracle(d.chi2, d.ld, d.gene_id, d.G, d.c)
starts = synth.jittered_start(G + 3, CHAINS, seed=5)
t0 = time.time()
trace, stats, n_evals = no.sample(co.logp_grad, starts, DRAWS, TUNE, seed=11)
print("oracle NUTS: %.1f s, %d logp+grad evals, mean depth %.2f, mean accept %.3f" % (
    time.time() - t0, n_evals, np.mean([s["depth"] for c in stats for s in c[TUNE:]]),
    np.mean([s["mean_tree_accept"] for c in stats for s in c[TUNE:]])))
D = G + 3
mean = trace.mean(axis=(0, 1))
sd = trace.std(axis=(0, 1))
ess = np.array([st.effective_n(trace[:, :, k]) for k in range(D)])
rhat = np.array([st.gelman_rubin(trace[:, :, k]) for k in range(D)])
print("min ess %.0f  max rhat %.4f" % (ess.min(), rhat.max()))
print("posterior mean e=%.4f W=%.4f mi=%.4f" % (mean[1], np.exp(trace[:, :, 0]).mean(), (1 / (1 + np.exp(-trace[:, :, 2]))).mean()))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "nuts_small.npz"),
                    M=M, G=G, seed=SEED, mean=mean, sd=sd, ess=ess, rhat=rhat,
                    depth_mean=np.mean([s["depth"] for c in stats for s in c[TUNE:]]),
                    accept_mean=np.mean([s["mean_tree_accept"] for c in stats for s in c[TUNE:]]))
