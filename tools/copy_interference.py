"""Does a concurrent PCIe copy slow the kernels?  Fused logp+grad kernel (and a plain torch FP64 kernel) timed
back-to-back while another stream runs pinned H2D / D2H copies."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from baghera_b200 import synth
from baghera_b200.engine import Engine
M, G, C = 1_100_000, 15_000, 64
d = synth.make_arrays(M, G)
eng = Engine(d.chi2, d.ld, d.gene_id, G, d.c, C)
q = eng.to_device_layout(synth.jittered_start(G + 3, C, seed=1))
logp = torch.empty(eng.Cp, dtype=torch.float64, device="cuda"); grad = torch.zeros_like(q)
n = 64 * 15003
h = torch.empty(n, dtype=torch.float64).pin_memory()
dbuf = torch.empty(n, dtype=torch.float64, device="cuda")
x = torch.randn(1 << 22, dtype=torch.float64, device="cuda")
cs = torch.cuda.Stream()


def timed(fn, copy, reps=300):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if copy:
        with torch.cuda.stream(cs):
            for _ in range(reps // 2):
                if copy == "d2h": h.copy_(dbuf, non_blocking=True)
                else: dbuf.copy_(h, non_blocking=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def k_fused(): eng.logp_grad(q, logp, grad)
def k_torch(): torch.addcmul(x, x, x, value=0.5, out=x)
for name, fn in (("fused logp+grad", k_fused), ("torch addcmul fp64 4M", k_torch)):
    for copy in (None, "h2d", "d2h", None):
        print("%-24s concurrent copy %-5s: %.1f us/launch" % (name, copy, timed(fn, copy)), flush=True)

# ---- the other direction: copy bandwidth while kernels run back-to-back on another stream -----------------
h2 = torch.empty(n, dtype=torch.float64).pin_memory()
d2 = torch.empty(n, dtype=torch.float64, device="cuda")
c1, c2, ks = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
y = torch.randn(1 << 24, dtype=torch.float32, device="cuda")
def k_sin(): torch.sin(y, out=y)


def copies(fn, R=200):
    torch.cuda.synchronize()
    if fn is not None:
        with torch.cuda.stream(ks):
            for _ in range(int(R * 172 / 40)): fn()
    t0 = time.perf_counter()
    for _ in range(R):
        with torch.cuda.stream(c1): dbuf.copy_(h, non_blocking=True)
        with torch.cuda.stream(c2): h2.copy_(d2, non_blocking=True)
    c1.synchronize(); c2.synchronize()
    dt = (time.perf_counter() - t0) / R
    torch.cuda.synchronize()
    return dt * 1e6


for name, fn in (("idle GPU", None), ("fused logp+grad", k_fused), ("torch addcmul fp64", k_torch), ("torch sin fp32 16M", k_sin), ("idle GPU", None)):
    print("h2d + d2h of 7.68 MB each while %-20s: %.1f us per pair" % (name, copies(fn)), flush=True)
