import time, sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from baghera_b200 import synth
from baghera_b200.engine import Engine
torch.cuda.init(); torch.zeros(1, device="cuda")
for M, G, model in ((1_100_000, 15_000, "gene"), (10_000_000, 1, "gw")):
    d = synth.make_arrays(M, G)
    for it in range(3):
        t0 = time.perf_counter()
        eng = Engine(d.chi2, d.ld, d.gene_id if model == "gene" else None, d.G, d.c, 64, model=model)
        t1 = time.perf_counter()
        eng.close()
    print("M=%d model=%s: Engine() incl. bgh_create + bgh_upload %.1f ms" % (M, model, (t1 - t0) * 1e3))
