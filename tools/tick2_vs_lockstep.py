"""Per-draw difference between the fused persistent sampler (bgh_tick2.cu) and the lock-step transition kernels."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
M, G, C = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (3000, 10, 4)))
tune, draws = (int(x) for x in (sys.argv[4:6] if len(sys.argv) > 5 else (0, 30)))
d = synth.make_arrays(M, G, seed=3)
eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, C)
print("split genes", eng.n_split_genes)
q0 = synth.jittered_start(G + 3, C, seed=1)
s = NutsSampler(eng, seed=5); s.init(q0)
a = s.sample_lockstep(draws=draws, tune=tune)
for v1 in ("0", "1"):
    os.environ["BGH_TICK"] = "1" if v1 == "1" else "2"
    s2 = NutsSampler(eng, seed=5); s2.init(q0)
    b = s2.sample(draws=draws, tune=tune, poll_ticks=16)
    ta, tb = a["trace"].cpu().numpy(), b["trace"].cpu().numpy()
    print("tick v%s: depths equal %s sizes equal %s" % ("1" if v1 == "1" else "2", np.array_equal(a["stats"][:, 0], b["stats"][:, 0]), np.array_equal(a["stats"][:, 2], b["stats"][:, 2])))
    rel = np.abs(ta - tb) / np.maximum(np.abs(ta), 1e-300)
    print("  max rel diff per draw:", " ".join("%.1e" % x for x in rel.reshape(draws, -1).max(axis=1)))
    print("  energy diff:", np.abs(a["stats"][:, 4] - b["stats"][:, 4]).max(), "step diff:", np.abs(a["stats"][:, 5] / b["stats"][:, 5] - 1).max())
    de = np.abs(a["stats"][:, 4] - b["stats"][:, 4])          # [transition, chain]
    print("  |energy diff| per transition (max over chains):", " ".join("%.0e" % x for x in de.max(axis=1)))
    print("  worst chain per transition:", " ".join(str(int(x)) for x in de.argmax(axis=1)))
eng.close()
