"""PCIe probe: pinned H2D / D2H bandwidth alone and concurrently, at the e2e step's transfer size."""
import time
import torch

n = 64 * 15003
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
R = 200


def run(h2d, d2h):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(R):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / R


for _ in range(2):
    for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
        dt = run(a, b)
        print("%s: %.1f us per 7.68 MB -> %.1f GB/s per direction" % (name, dt * 1e6, n * 8 / dt / 1e9))
