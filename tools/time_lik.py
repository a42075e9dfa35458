"""Times lik_kernel / finalize_kernel (library CUDA events) for plan-parameter sweeps."""
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
ap.add_argument("--C", type=int, nargs="+", default=[64])
ap.add_argument("--kappa", type=float, nargs="+", default=[6.0])
ap.add_argument("--n", type=int, default=300)
ap.add_argument("--model", default="gene")
ap.add_argument("--kappa-warp", type=float, nargs="+", default=[None])
args = ap.parse_args()
d = synth.make_arrays(args.M, args.G)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for C in args.C:
    for kappa, kw in [(k, w) for k in args.kappa for w in args.kappa_warp]:
        os.environ["BGH_PLAN_KAPPA"] = str(kappa)
        if kw is not None:
            os.environ["BGH_PLAN_KAPPA_WARP"] = str(kw)
        eng = Engine(d.chi2, d.ld, d.gene_id if args.model == "gene" else None, d.G, d.c, C, model=args.model)
        Q = synth.jittered_start(eng.D, C, seed=1, model=args.model)
        q = eng.to_device_layout(Q)
        lp = torch.empty(eng.Cp, dtype=torch.float64, device="cuda")
        g = torch.empty_like(q)
        for _ in range(5):
            eng.logp_grad(q, lp, g)
        torch.cuda.synchronize()
        res = {}
        for mode in ("flushed", "steady"):
            eng.set_profiling(True)
            for i in range(args.n):
                if mode == "flushed":
                    flush.fill_(i & 255)
                eng.logp_grad(q, lp, g)
            torch.cuda.synchronize()
            a, b, n = eng.kernel_time_ms(reset=True)
            eng.set_profiling(False)
            res[mode] = (a * 1e3, b * 1e3)
        print("C=%d kappa=%.1f kappa_warp=%s split=%d  flushed: lik %.2f us fin %.2f us | steady: lik %.2f us fin %.2f us" % (
            C, kappa, kw, eng.n_split_genes, res["flushed"][0], res["flushed"][1], res["steady"][0], res["steady"][1]), flush=True)
        eng.close()
