import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
d = synth.make_arrays(1_100_000, 15_000)
eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, 64)
for poll, ng in ((64, '0'), (64, '1')):
    os.environ['BGH_NUTS_NO_GRAPH'] = ng
    s = NutsSampler(eng, seed=1)
    s.init(synth.jittered_start(eng.D, 64, seed=2))
    out = s.sample(0, 40, poll_ticks=poll)
    print("nograph=%s" % ng, "async poll=%d ticks %d secs %.3f -> %.1f us/tick" % (poll, out["ticks"], out["seconds_total"], out["seconds_total"] / out["ticks"] * 1e6), flush=True)
s = NutsSampler(eng, seed=1)
s.init(synth.jittered_start(eng.D, 64, seed=2))
out = s.sample_lockstep(0, 40)
n = out["leapfrogs"].sum()
print("lockstep leaves %d secs %.3f -> %.1f us/leaf" % (n, out["seconds_tune"], out["seconds_tune"] / n * 1e6))
