"""Minimal driver for profiling: a few batched logp+grad evaluations at a BASELINE.json shape."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1_100_000)
ap.add_argument("--G", type=int, default=15_000)
ap.add_argument("--C", type=int, default=64)
ap.add_argument("--n", type=int, default=6)
ap.add_argument("--model", default="gene")
ap.add_argument("--sharded", action="store_true", help="the N > 1 kernel (logp_grad_chainx_kernel) on a one-rank sharded context")
args = ap.parse_args()
d = synth.make_arrays(args.M, args.G)
shard = None
if args.sharded:
    from baghera_b200.sharding import shard_dataset
    shard = shard_dataset(d.chi2, d.ld, d.gene_id, d.G, 0, 1)
eng = Engine(d.chi2, d.ld, d.gene_id if args.model == "gene" else None, d.G, d.c, args.C, model=args.model,
             shard=shard.cfg if shard else None)
D = eng.D
Q = synth.jittered_start(D, args.C, seed=1, model=args.model)
q = eng.to_device_layout(Q)
if args.sharded:
    eng.peer_export()
    eng.peer_connect_local(0, [eng])
lp = torch.empty(eng.Cp, dtype=torch.float64, device="cuda")
g = torch.empty_like(q)
for _ in range(args.n):
    if args.sharded:
        eng.logp_grad_sharded(q, lp, g)
    else:
        eng.logp_grad(q, lp, g)
torch.cuda.synchronize()
print("ok", float(lp[0]))
