"""Short async NUTS run at a BASELINE.json shape (for profiling)."""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1_100_000); ap.add_argument("--G", type=int, default=15_000)
ap.add_argument("--C", type=int, default=64); ap.add_argument("--tune", type=int, default=12); ap.add_argument("--draws", type=int, default=0)
ap.add_argument("--lockstep", action="store_true")
a = ap.parse_args()
d = synth.make_arrays(a.M, a.G)
eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, a.C)
s = NutsSampler(eng, seed=1)
s.init(synth.jittered_start(eng.D, a.C, seed=2))
out = (s.sample_lockstep if a.lockstep else s.sample)(a.draws, a.tune)
print("ok ticks", out.get("ticks"), "leapfrogs", float(np.sum(out["leapfrogs"])), "secs", out["seconds_tune"] + out["seconds_draws"])
