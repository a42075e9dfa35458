"""Per-phase timing of the persistent NUTS tick kernel (globaltimer stamps of every CTA, first 16 ticks
of the LAST launch)."""
import argparse, ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
from baghera_b200._lib import check
ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1_100_000); ap.add_argument("--G", type=int, default=15_000)
ap.add_argument("--C", type=int, default=64); ap.add_argument("--tune", type=int, default=30)
a = ap.parse_args()
d = synth.make_arrays(a.M, a.G)
eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, a.C)
check(eng.L.bgh_debug_trace(eng.ctx, 1, None, 0), eng.ctx)
s = NutsSampler(eng, seed=1)
s.init(synth.jittered_start(eng.D, a.C, seed=2))
out = s.sample(0, a.tune, poll_ticks=64)
print("ticks %d secs %.3f -> %.1f us/tick" % (out["ticks"], out["seconds_total"], out["seconds_total"] / out["ticks"] * 1e6))
n_cta = 148
words = eng.Cp // 64 * n_cta * 16 * 8 if eng.Cp >= 64 else n_cta * 16 * 8
buf = np.zeros(words, dtype=np.uint64)
check(eng.L.bgh_debug_trace(eng.ctx, 1, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), words), eng.ctx)
t = buf[:8 * n_cta * 16].reshape(8, n_cta, 16).astype(np.int64)
names = ["pre", "barrier1", "lik", "barrier2+fin", "barrier3+post", "barrier4", "scalar", "barrier5"]
for tick in (1, 3, 6):
    tt = t[tick] - t[tick].min()
    print("tick", tick)
    for k in range(8):
        nxt = t[tick, :, k + 1] if k < 7 else t[tick + 1, :, 0]
        dur = (nxt - t[tick, :, k]) / 1e3
        print("  %-14s start p50 %7.2f  dur min %6.2f p50 %6.2f max %6.2f us" % (names[k], np.median(tt[:, k]) / 1e3, dur.min(), np.median(dur), dur.max()))
    print("  tick total %.2f us" % ((t[tick + 1, 0, 0] - t[tick, 0, 0]) / 1e3))
    sub = ["post: barrier3 + Z refill", "post: core rows", "post: level dots", "post: CTA reduce"]
    marks = [t[tick, :, 4], t[tick, :, 8], t[tick, :, 9], t[tick, :, 10], t[tick, :, 5]]
    for k in range(4):
        dur = (marks[k + 1] - marks[k]) / 1e3
        print("    %-26s dur min %6.2f p50 %6.2f max %6.2f us" % (sub[k], dur.min(), np.median(dur), dur.max()))
eng.close()
