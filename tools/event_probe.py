"""Cost of stream-ordered event records / waits between pinned copies (device timeline bubbles)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from baghera_b200 import synth
from baghera_b200.engine import Engine
dd = synth.make_arrays(1_100_000, 15_000)
eng = Engine(dd.chi2, dd.ld, dd.gene_id, 15_000, dd.c, 64)
q_dc = eng.to_device_layout(synth.jittered_start(15_003, 64, seed=1))
logp = torch.empty(eng.Cp, dtype=torch.float64, device="cuda"); grad = torch.zeros_like(q_dc)
q_cd = torch.zeros((64, 15_003), dtype=torch.float64, device="cuda")

n = 64 * 15003
R = 200
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2, s3 = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
s1b, s2b = torch.cuda.Stream(), torch.cuda.Stream()
tiny = torch.zeros(32, device="cuda")


def run(mode):
    evs_up = [torch.cuda.Event() for _ in range(4)]
    evs_c = [torch.cuda.Event() for _ in range(4)]
    evs_dn = [torch.cuda.Event() for _ in range(4)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(R):
        b = i % 4
        u, d = (s1, s2) if (not mode.startswith("two") or i % 2 == 0) else (s1b, s2b)
        if mode in ("wait-old", "full", "two-streams", "two+kernels") and i >= 4:
            u.wait_event(evs_c[b]); s3.wait_event(evs_dn[b])
        with torch.cuda.stream(u):
            d_in.copy_(h_in, non_blocking=True)
            if mode != "none": evs_up[b].record(u)
        if mode in ("full", "two-streams", "two+kernels", "two+kernels-nowait"):
            s3.wait_event(evs_up[b])
            with torch.cuda.stream(s3):
                if mode.startswith("two+kernels"):
                    qq = eng.to_device_layout(q_cd)
                    eng.logp_grad(qq, logp, grad)
                    eng.from_device_layout(grad)
                else:
                    tiny.add_(1.0)
                evs_c[b].record(s3)
            d.wait_event(evs_c[b])
        elif mode == "wait-old":
            evs_c[b].record(s3)
        with torch.cuda.stream(d):
            h_out.copy_(d_out, non_blocking=True)
            if mode != "none": evs_dn[b].record(d)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / R * 1e6


for mode in ("none", "none", "full", "two-streams", "two+kernels", "two+kernels-nowait", "none"):
    print("%-20s %.1f us per h2d+d2h pair" % (mode, run(mode)), flush=True)
