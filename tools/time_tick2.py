"""Tick time of the fused persistent sampler (bgh_tick2.cu) against nuts_tick_kernel (BGH_TICK_V1=1), with the
phase stamps of CTA 0 (likelihood pass | barrier | scalar phase | barrier)."""
import argparse, ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
from baghera_b200._lib import check
ap = argparse.ArgumentParser()
ap.add_argument("--M", type=int, default=1_100_000); ap.add_argument("--G", type=int, default=15_000)
ap.add_argument("--C", type=int, default=64); ap.add_argument("--tune", type=int, default=40)
ap.add_argument("--v1", action="store_true", help="also time the round-1 tick kernel")
ap.add_argument("--cta", default="0", help="comma list of CTAs whose phase stamps are printed")
a = ap.parse_args()
d = synth.make_arrays(a.M, a.G)
world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
shard = None
if world > 1:      # torchrun: the same chains on gene-block shards (fused kernel only)
    import torch.distributed as dist
    from baghera_b200.sharding import shard_dataset, connect_peers
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    shard = shard_dataset(d.chi2, d.ld, d.gene_id, d.G, rank, world)
    eng = Engine(shard.chi2, shard.ld, shard.gene_id, shard.n_genes, d.c, a.C, device=lr, shard=shard.cfg)
    lc = torch.tensor([eng.lik_const], dtype=torch.float64, device="cuda")
    dist.all_reduce(lc)
    eng.set_lik_const(float(lc.item()))
    connect_peers(eng, dist, torch.device("cuda", lr))
    if rank != 0:
        sys.stdout = open(os.devnull, "w")
else:
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, a.C)
print("world", world, "split genes (evaluation plan)", eng.n_split_genes, "local rows", eng.D, flush=True)
check(eng.L.bgh_debug_trace(eng.ctx, 1, None, 0), eng.ctx)
runs = [("2", m) for m in a.cta.split(",")] + ([("1", "0")] if a.v1 else [])
for ver, cta in runs:
    os.environ["BGH_TICK"] = ver
    os.environ["BGH_T2_TRACE_CTA"] = cta
    print("--- stamps of CTA", cta)
    if shard is not None:
        s = NutsSampler(eng, seed=1, row_coords=shard.row_coords(), n_dim_total=d.G + 3)
        s.init(shard.local_q(synth.jittered_start(d.G + 3, a.C, seed=2)))
        dist.barrier()
    else:
        s = NutsSampler(eng, seed=1)
        s.init(synth.jittered_start(eng.D, a.C, seed=2))
    out = s.sample(0, a.tune, poll_ticks=64)
    st = out["stats"]
    print("tick v%s: ticks %d secs %.3f -> %.1f us/tick; leapfrogs/transition %.1f accept %.3f" % (
        ver, out["ticks"], out["seconds_total"], out["seconds_total"] / out["ticks"] * 1e6, st[:, 2].mean(), st[:, 1].mean()), flush=True)
    if ver == "2":
        words = max(eng.Cp // 64, 1) * 148 * 16 * 8
        buf = np.zeros(words, dtype=np.uint64)
        check(eng.L.bgh_debug_trace(eng.ctx, 1, buf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), words), eng.ctx)
        t = buf[:64 * 16].reshape(64, 16).astype(np.int64)
        n = min(out["ticks"], 63)
        if world > 1:
            order = [0, 10, 8, 9, 1, 2, 5, 11, 12, 6, 7, 3, 4]
            names = ["rare events", "pre (half step)", "likelihood pass", "post (second half + sums)", "barrier 1", "S: reduce partials + slots",
                     "S: local late rows", "S: peer exchange (incl. rank skew)", "S: closing + late finish", "S: state machine",
                     "S: late enter + store", "barrier 2"]
        else:
            order = [0, 10, 8, 9, 1, 2, 5, 6, 7, 3, 4]
            names = ["rare events", "pre (half step)", "likelihood pass", "post (second half + sums)", "barrier 1", "S: reduce partials + slots",
                     "S: closing + late finish", "S: state machine", "S: late enter + store", "barrier 2"]
        for k, name in enumerate(names):
            dur = (t[2:n, order[k + 1]] - t[2:n, order[k]]) / 1e3
            print("  %-36s p50 %6.2f  mean %6.2f  max %6.2f us" % (name, np.median(dur), dur.mean(), dur.max()))
        print("  tick total p50 %.2f us" % np.median(np.diff(t[2:n, 0]) / 1e3))
        if os.environ.get("BGH_T2_RAW"):
            for k in range(2, 6):
                print("   raw tick", k, [int(x - t[k, 0]) if x > 1e15 else int(x) for x in t[k]])
        if os.environ.get("BGH_T2_DETAIL"):
            print("  own rows", t[2, 15], " per tick: n_rare cmd_us items_us(thread0) sync_us rest_us")
            for k in range(2, min(n, 40)):
                print("   %3d  n_rare %2d  cmd %6.2f  t0-items %6.2f  wait %6.2f  sums %6.2f" % (k, t[k, 14], (t[k, 11] - t[k, 0]) / 1e3,
                      (t[k, 12] - t[k, 11]) / 1e3 if t[k, 14] else 0, (t[k, 13] - t[k, 12]) / 1e3 if t[k, 14] else 0, (t[k, 10] - t[k, 13]) / 1e3 if t[k, 14] else 0))
eng.close()
if world > 1:
    dist.destroy_process_group()
