import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
d = synth.make_arrays(1_100_000, 15_000)
eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, 64)
q = eng.to_device_layout(synth.jittered_start(eng.D, 64, seed=1))
g = torch.empty_like(q); lp = torch.empty(eng.Cp, dtype=torch.float64, device="cuda")
payload = torch.zeros((eng.payload_rows, eng.Cp), dtype=torch.float64, device="cuda")
for _ in range(12):
    eng.logp_grad_local(q, g, payload)
    eng.logp_grad_finish(q, payload, lp, g)
torch.cuda.synchronize(); print("ok")
