"""CPU: the committed golden vectors are reproduced by every oracle form (incl. the C restatement)."""
import os

import numpy as np

from baghera_b200 import synth
from oracle import model_oracle as mo
from oracle.c_oracle import COracle
from helpers import grad_rel_err


def test_golden_vectors_oracle():
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "logp_grad_cfg1.npz"))
    d = synth.make_arrays(int(z["M"]), int(z["G"]), seed=int(z["seed"]))
    chi2, ld, cat = d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64)
    ss = mo.SuffStats(chi2, ld, cat, d.G)
    for fix, key in ((False, "free"), (True, "fix")):
        co = COracle(chi2, ld, cat, d.G, d.c, fix)
        lp_c, g_c = co.logp_grad_batch(z["Q"], n_threads=2)
        for i, q in enumerate(z["Q"]):
            lp_s, g_s = mo.logp_grad_suffstat(q, ss, d.c, fix)
            for lp, g in ((lp_c[i], g_c[i]), (lp_s, g_s)):
                assert abs(lp - z["logp_" + key][i]) <= 1e-11 * abs(lp)
                assert grad_rel_err(g, z["grad_" + key][i]) <= 1e-9
