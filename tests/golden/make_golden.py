"""Generates tests/golden/logp_grad_cfg1.npz from the oracle (run once; the .npz is committed).

The reference cannot be imported here (pymc3==3.6 / theano are not installable), so the
golden vectors freeze the *oracle's* closed form after it has been cross-checked against the
PyMC3-shaped autograd form, the sufficient-statistics form and the C restatement
(tests/test_oracle.py).  Parity against PyMC3 itself stays unpinned.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from baghera_b200 import synth                    # noqa: E402
from oracle import model_oracle as mo             # noqa: E402

M, G, SEED = 20_000, 200, 20211
d = synth.make_arrays(M, G, seed=SEED)
chi2, ld, cat = d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64)
rng = np.random.default_rng(77)
Q = synth.jittered_start(G + 3, 6, seed=77)
Q[4, 0], Q[4, 2] = -20.0, 20.0          # extreme transforms
Q[5, 0], Q[5, 2] = 20.0, -20.0
Q[3, 3:] = rng.normal(-0.5, 1.0, G)     # negative betas
out = dict(M=M, G=G, seed=SEED, Q=Q)
for fix, key in ((False, "free"), (True, "fix")):
    lp, g = mo.batch_closed(Q, chi2, ld, cat, d.c, fix)
    for i in range(Q.shape[0]):           # the autograd form must agree before freezing
        la, ga = mo.logp_grad_autograd(Q[i], chi2, ld, cat, d.c, fix)
        assert abs(la - lp[i]) <= 1e-12 * abs(la)
        assert np.max(np.abs(ga - g[i])) <= 1e-9 * np.max(np.abs(ga))
    out["logp_" + key], out["grad_" + key] = lp, g
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "logp_grad_cfg1.npz"), **out)
print("wrote golden vectors", {k: np.shape(v) for k, v in out.items()})
