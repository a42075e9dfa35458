"""Stage timeline of the host-entry pipeline (BGH_E2E_TRACE=1): where an e2e step spends its time."""
import os, sys, time
os.environ.setdefault("BGH_E2E_TRACE", "1")   # 1: device stage events + host enqueue times, 2: host times only
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from baghera_b200 import synth
from baghera_b200.engine import Engine
M, G, C = 1_100_000, 15_000, 64
d = synth.make_arrays(M, G)
eng = Engine(d.chi2, d.ld, d.gene_id, G, d.c, C)
Q = synth.jittered_start(G + 3, C, seed=1)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
qh = torch.empty((n, C, G + 3), dtype=torch.float64).pin_memory(); qh.copy_(torch.from_numpy(np.broadcast_to(Q, (n, C, G + 3)).copy()))
gh = torch.empty((n, C, G + 3), dtype=torch.float64).pin_memory()
lh = torch.empty((n, C), dtype=torch.float64).pin_memory()
eng.logp_grad_host_ptr(4, qh.data_ptr(), lh.data_ptr(), gh.data_ptr())
for rep in range(int(os.environ.get("REPS", "1"))):
    t0 = time.perf_counter()
    eng.logp_grad_host_ptr(n, qh.data_ptr(), lh.data_ptr(), gh.data_ptr())
    print("wall %.1f us/step" % ((time.perf_counter() - t0) / n * 1e6), flush=True)
