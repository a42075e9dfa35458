"""Per-warp stamps of the likelihood pass INSIDE the fused sampler kernel (last tick of the run; BGH_T2_LIK_TRACE=1):
where does the pass spend its time on a shard?  torchrun for gene-block shards, plain python for one GPU."""
import argparse, ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BGH_T2_LIK_TRACE"] = "1"
os.environ["BGH_TICK"] = "2"
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
from baghera_b200._lib import check
ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1_100_000); ap.add_argument("--G", type=int, default=15_000)
ap.add_argument("--C", type=int, default=64); ap.add_argument("--tune", type=int, default=6)
a = ap.parse_args()
d = synth.make_arrays(a.M, a.G)
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
if world > 1:
    import torch.distributed as dist
    from baghera_b200.sharding import shard_dataset, connect_peers
    lr = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    shard = shard_dataset(d.chi2, d.ld, d.gene_id, d.G, rank, world)
    eng = Engine(shard.chi2, shard.ld, shard.gene_id, shard.n_genes, d.c, a.C, device=lr, shard=shard.cfg)
    lc = torch.tensor([eng.lik_const], dtype=torch.float64, device="cuda"); dist.all_reduce(lc); eng.set_lik_const(float(lc.item()))
    connect_peers(eng, dist, torch.device("cuda", lr))
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
    s = NutsSampler(eng, seed=1, row_coords=shard.row_coords(), n_dim_total=d.G + 3)
    check(eng.L.bgh_debug_trace(eng.ctx, 1, None, 0), eng.ctx)
    s.init(shard.local_q(synth.jittered_start(d.G + 3, a.C, seed=2)))
    dist.barrier()
else:
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, a.C)
    s = NutsSampler(eng, seed=1)
    check(eng.L.bgh_debug_trace(eng.ctx, 1, None, 0), eng.ctx)
    s.init(synth.jittered_start(eng.D, a.C, seed=2))
out = s.sample(0, a.tune, poll_ticks=64)
words = max(eng.Cp // 64, 1) * 148 * 16 * 8
buf = np.zeros(words, dtype=np.uint64)
check(eng.L.bgh_debug_trace(eng.ctx, 1, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), words), eng.ctx)
t = buf.reshape(-1, 8).astype(np.int64)[:148 * 16]
t0 = t[:, 0].min()
enter, setup, stream, bar, exit_ = [(t[:, k] - t0) / 1e3 for k in range(5)]
print("world %d: ticks %d, %.1f us/tick; likelihood pass of the last tick, %d warps" % (world, out["ticks"], out["seconds_total"] / out["ticks"] * 1e6, t.shape[0]))
for name, x in (("enter", enter), ("setup done", setup), ("stream done", stream), ("after CTA sync", bar), ("exit", exit_)):
    print("  %-15s min %.2f  p50 %.2f  p90 %.2f  max %.2f us" % (name, x.min(), np.median(x), np.percentile(x, 90), x.max()))
print("  setup duration   p50 %.2f max %.2f us" % (np.median(setup - enter), (setup - enter).max()))
dur = stream - setup
print("  stream duration  min %.2f p50 %.2f p90 %.2f max %.2f us" % (dur.min(), np.median(dur), np.percentile(dur, 90), dur.max()))
pc = (exit_.reshape(-1, 16).max(1) - enter.reshape(-1, 16).min(1))
print("  per-CTA pass duration: min %.2f p50 %.2f max %.2f us" % (pc.min(), np.median(pc), pc.max()))
ce = enter.reshape(-1, 16).min(1)
print("  per-CTA enter (PRE done): min %.2f p50 %.2f max %.2f us" % (ce.min(), np.median(ce), ce.max()))
for c in (0, 40, 60, 100, 147):
    print("   CTA %3d: enter %.2f setup %.2f stream-end min %.2f max %.2f exit %.2f" % (c, enter[c*16:(c+1)*16].min(), setup[c*16:(c+1)*16].max(), stream[c*16:(c+1)*16].min(), stream[c*16:(c+1)*16].max(), exit_[c*16:(c+1)*16].max()))
off = eng.partition(sampler=True)
_, _, _, gid_dev = eng.device_stream()
for c in (40, 60):
    print("   CTA %d warps: (SNPs, genes, stream us)" % c, [(int(off[v + 1] - off[v]), int(len(np.unique(gid_dev[off[v]:off[v + 1]]))), round(float(dur[v]), 1)) for v in range(c * 16, c * 16 + 16)])
eng.close()
if world > 1:
    dist.destroy_process_group()
