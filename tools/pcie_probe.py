"""PCIe probe: pinned H2D / D2H bandwidth alone and concurrently, at the e2e step's transfer size, as a
function of the host-side footprint (a ring of nbuf distinct pinned buffers per direction)."""
import sys
import time
import torch

n = 64 * 15003
R = 256
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")


def run(h_in, h_out, h2d, d2h):
    nb = h_in.shape[0]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(R):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in[i % nb], non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out[i % nb].copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / R


for nbuf in (1, 4, 16, 64):
    h_in = torch.empty((nbuf, n), dtype=torch.float64).pin_memory()
    h_out = torch.empty((nbuf, n), dtype=torch.float64).pin_memory()
    h_in.fill_(1.0); h_out.fill_(0.0)
    run(h_in, h_out, 1, 1)
    res = []
    for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
        dt = run(h_in, h_out, a, b)
        res.append("%s %.1f us (%.1f GB/s)" % (name, dt * 1e6, n * 8 / dt / 1e9))
    print("ring of %2d x 7.68 MB per direction: %s" % (nbuf, ", ".join(res)), flush=True)
    del h_in, h_out
