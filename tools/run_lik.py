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
args = ap.parse_args()
d = synth.make_arrays(args.M, args.G)
eng = Engine(d.chi2, d.ld, d.gene_id if args.model == "gene" else None, d.G, d.c, args.C, model=args.model)
D = eng.D
Q = synth.jittered_start(D, args.C, seed=1, model=args.model)
q = eng.to_device_layout(Q)
for _ in range(args.n):
    lp, g = eng.logp_grad(q)
torch.cuda.synchronize()
print("ok", float(lp[0]))
