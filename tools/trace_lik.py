"""Per-warp phase timing of lik_kernel (needs a GPU): where does the kernel's fixed overhead go?"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200._lib import check
from baghera_b200.engine import Engine

ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1_100_000)
ap.add_argument("--G", type=int, default=15_000)
ap.add_argument("--C", type=int, default=64)
ap.add_argument("--flush", action="store_true")
args = ap.parse_args()
d = synth.make_arrays(args.M, args.G)
eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, args.C)
q = eng.to_device_layout(synth.jittered_start(eng.D, args.C, seed=1))
lp = torch.empty(eng.Cp, dtype=torch.float64, device="cuda")
g = torch.empty_like(q)
check(eng.L.bgh_debug_trace(eng.ctx, 1, None, 0), eng.ctx)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
n_cta = int(eng.L.bgh_n_groups(eng.ctx)) // 16
for it in range(4):
    if args.flush:
        flush.fill_(it)
    eng.logp_grad(q, lp, g)
torch.cuda.synchronize()
words = (eng.Cp // (64 if args.C > 32 else eng.Cp)) * n_cta * 16 * 8
buf = np.zeros(words, dtype=np.uint64)
check(eng.L.bgh_debug_trace(eng.ctx, 1, buf.ctypes.data_as(ctypes.c_void_p), words), eng.ctx)
t = buf.reshape(-1, 8).astype(np.int64)
t0 = t[:, 0].min()
enter, setup, stream, bar, exit_ = [(t[:, k] - t0) / 1e3 for k in range(5)]
print("warps", t.shape[0], "kernel span (first enter -> last exit) %.2f us" % exit_.max())
for name, x in (("enter", enter), ("setup done", setup), ("stream done", stream), ("barrier passed", bar), ("exit", exit_)):
    print("  %-15s min %.2f  p50 %.2f  p90 %.2f  max %.2f us" % (name, x.min(), np.median(x), np.percentile(x, 90), x.max()))
print("  setup duration   p50 %.2f max %.2f us" % (np.median(setup - enter), (setup - enter).max()))
dur = stream - setup
print("  stream duration  min %.2f p50 %.2f p90 %.2f max %.2f us" % (dur.min(), np.median(dur), np.percentile(dur, 90), dur.max()))
cta_end = stream.reshape(-1, 16).max(axis=1)
print("  per-CTA stream end: min %.2f p50 %.2f max %.2f us" % (cta_end.min(), np.median(cta_end), cta_end.max()))
print("  within-CTA spread of stream end (max-min): p50 %.2f max %.2f us" % (
    np.median(stream.reshape(-1, 16).max(1) - stream.reshape(-1, 16).min(1)),
    (stream.reshape(-1, 16).max(1) - stream.reshape(-1, 16).min(1)).max()))
off = eng.partition()
lens = np.diff(off)
print("  SNPs per group: min %d p50 %d max %d" % (lens.min(), np.median(lens), lens.max()))
# slowest CTAs: exclusive or mixed?
order = np.argsort(-cta_end)[:8]
print("  slowest CTAs:", [(int(c), round(float(cta_end[c]), 2), int(lens[c * 16:(c + 1) * 16].sum())) for c in order])
order = np.argsort(cta_end)[:8]
print("  fastest CTAs:", [(int(c), round(float(cta_end[c]), 2), int(lens[c * 16:(c + 1) * 16].sum())) for c in order])
eng.close()
