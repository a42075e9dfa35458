"""Chains as independent replicas over several GPUs of one box.

``pm.sample(chains=N, cores=K)`` of the reference (ref: baghera_tool/gene_regression.py:100-106) runs its chains as
independent processes; here the chains of one analysis are dealt in contiguous blocks to the listed devices, every
device runs its block through its own Engine + persistent NUTS kernel (one host thread per device: the C calls
release the GIL), and the traces are gathered on the first device over NVLink.  No data-path collective.
"""
from __future__ import annotations

import threading

import numpy as np


def sample_on_devices(make_engine, start, draws, tune, devices, seed, trace_dtype=None, progress=None, target_accept=0.90,
                      stream_stats=False):
    """make_engine(device, n_chains) -> Engine; start: float64[C, D] start points of ALL chains.
    Returns (out, launches): out as NutsSampler.sample (trace [draws, D, C] on devices[0], stats [n, 8, C], ...)."""
    import torch

    from .sampler import NutsSampler

    devices = [int(d) for d in devices]
    C = int(start.shape[0])
    blocks = [b for b in np.array_split(np.arange(C), min(len(devices), C)) if len(b)]
    devices = devices[:len(blocks)]
    results, errors = [None] * len(blocks), [None] * len(blocks)

    def work(i):
        try:
            dev, idx = devices[i], blocks[i]
            torch.cuda.set_device(dev)
            eng = make_engine(dev, len(idx))
            try:
                # Philox keys by the GLOBAL chain index and the start mean over ALL chains: the draws do not depend on the
                # device list (--devices 0,1 reproduces --device 0)
                smp = NutsSampler(eng, seed=seed, target_accept=target_accept, chain_offset=int(idx[0]))
                smp.init(start[idx], start_mean=start.mean(axis=0))
                out = smp.sample(int(draws), int(tune), trace_dtype=trace_dtype, progress=progress if i == 0 else None,
                                 stream_stats=stream_stats)
                torch.cuda.synchronize(dev)
                results[i] = (out, eng.launch_count)
            finally:
                eng.close()             # the trace lives in torch memory, the engine can go
        except BaseException as e:      # re-raised in the caller's thread
            errors[i] = e

    if len(blocks) == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(i,), name="baghera-dev%d" % devices[i]) for i in range(len(blocks))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    for e in errors:
        if e is not None:
            raise e
    outs = [r[0] for r in results]
    dev0 = torch.device("cuda", devices[0])
    merged = dict(outs[0])
    if len(outs) > 1:
        merged["trace"] = torch.cat([o["trace"].to(dev0, non_blocking=True) for o in outs], dim=2)
        merged["stats"] = np.concatenate([o["stats"] for o in outs], axis=2)
        if stream_stats:
            merged["k6"] = torch.cat([o["k6"].to(dev0, non_blocking=True) for o in outs], dim=2)
            merged["head_trace"] = torch.cat([o["head_trace"].to(dev0, non_blocking=True) for o in outs], dim=2)
        merged["leapfrogs"] = np.mean([o["leapfrogs"] for o in outs], axis=0)
        merged["ticks"] = int(max(o["ticks"] for o in outs))
        for k in ("seconds_tune", "seconds_draws", "seconds_total"):
            merged[k] = float(max(o[k] for o in outs))
        torch.cuda.synchronize(dev0)
    merged["devices"] = devices
    return merged, int(sum(r[1] for r in results))
