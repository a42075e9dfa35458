import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
C,M,G=100,20000,300
d=synth.make_arrays(M,G,seed=5)
eng=Engine(d.chi2,d.ld,d.gene_id,d.G,d.c,C)
q0=synth.jittered_start(G+3,C,seed=2)
s=NutsSampler(eng,seed=11); s.init(q0); a=s.sample_lockstep(draws=6,tune=14)
s2=NutsSampler(eng,seed=11); s2.init(q0); b=s2.sample(draws=6,tune=14,poll_ticks=32)
ta,tb=a["trace"].cpu().numpy(),b["trace"].cpu().numpy()
rel=np.abs(ta-tb)/np.maximum(np.abs(ta),1e-300)
print("max rel",rel.max(), "where", np.unravel_index(rel.argmax(), rel.shape))
print("per-draw max", rel.max(axis=(1,2)))
print("per-chain max (top5)", np.sort(rel.max(axis=(0,1)))[-5:], np.argsort(rel.max(axis=(0,1)))[-5:])
print("stats eq", [np.array_equal(a["stats"][:,k], b["stats"][:,k]) for k in range(8)])
print("eps diff", np.abs(a["stats"][:,5]-b["stats"][:,5]).max())
