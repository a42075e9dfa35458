"""torchrun --nproc-per-node N tools/sharded_sampler_check.py : the gene-block sharded NUTS sampler (every rank
owns its rows of q; one LL-protocol peer exchange of the per-chain sums per leapfrog inside the persistent kernel,
DRIVER=persistent, or the lock-step transition kernels, DRIVER=lockstep) against the single-GPU lock-step sampler
run with the same seed on rank 0."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sampler import NutsSampler
from baghera_b200.sharding import shard_dataset, shard_snps_only, connect_peers

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
M, G, C = int(os.environ.get("M", 20_000)), int(os.environ.get("G", 200)), int(os.environ.get("C", 8))
TUNE, DRAWS = int(os.environ.get("TUNE", 12)), int(os.environ.get("DRAWS", 6))
model = os.environ.get("MODEL", "gene")
gw = model == "gw"
d = synth.make_arrays(M, G)
D = 2 if gw else G + 3
Q0 = synth.jittered_start(D, C, seed=2, model=model)
driver = os.environ.get("DRIVER", "persistent")
sh = shard_snps_only(d.chi2, d.ld, rank, world) if gw else shard_dataset(d.chi2, d.ld, d.gene_id, G, rank, world)
eng = Engine(sh.chi2, sh.ld, None if gw else sh.gene_id, sh.n_genes, d.c, C, model=model, device=lr, shard=sh.cfg)
lc = torch.tensor([eng.lik_const], dtype=torch.float64, device=dev)
dist.all_reduce(lc)
eng.set_lik_const(float(lc.item()))
connect_peers(eng, dist, dev)
s = NutsSampler(eng, seed=11, row_coords=None if gw else sh.row_coords(), n_dim_total=D)
s.init(Q0 if gw else sh.local_q(Q0))
dist.barrier()
t0 = time.perf_counter()
out = s.sample_lockstep(DRAWS, TUNE) if driver == "lockstep" else s.sample(DRAWS, TUNE)
secs = time.perf_counter() - t0
if driver != "lockstep":
    out["leapfrogs"] = out["stats"][:, 2].sum(axis=1)        # per-transition tree sizes summed over chains
assert eng.peer_status() == 0, eng.peer_status()
st = torch.as_tensor(out["stats"], device=dev)
allst = [torch.empty_like(st) for _ in range(world)]
dist.all_gather(allst, st)
same = all(torch.equal(allst[0], x) for x in allst)            # every rank took bit-identical decisions
tr = out["trace"].cpu().numpy()                                 # [draws, D_local, C]
ok = True
if rank == 0:
    print("sharded x%d (%s): %d transitions, %d leapfrogs%s, %.2f s; stats bit-identical on all ranks: %s"
          % (world, driver, TUNE + DRAWS, int(out["leapfrogs"].sum()), (", %d ticks = %.1f us/tick" % (out["ticks"], secs / out["ticks"] * 1e6))
             if driver != "lockstep" else "", secs, same), flush=True)
gathered = [None] * world
dist.all_gather_object(gathered, (sh.row_coords(2 if gw else 3), tr))
if rank == 0:
    full = np.full((DRAWS, D, C), np.nan)
    for rc, t in gathered:
        if gw:
            full[:] = t
        else:
            full[:, rc] = t
    e1 = Engine(d.chi2, d.ld, None if gw else d.gene_id, G, d.c, C, model=model, device=lr)
    s1 = NutsSampler(e1, seed=11)
    s1.init(Q0)
    ref = s1.sample_lockstep(DRAWS, TUNE)
    rt = ref["trace"].cpu().numpy()
    dep = np.array_equal(ref["stats"][:, 0], out["stats"][:, 0]) and np.array_equal(ref["stats"][:, 2], out["stats"][:, 2])
    err = np.max(np.abs(full - rt) / np.maximum(np.abs(rt), 1e-3))
    eps = np.max(np.abs(ref["stats"][:, 5] - out["stats"][:, 5]) / ref["stats"][:, 5])
    print("vs single GPU: tree depths and sizes equal: %s ; step size rel diff %.2e ; positions rel diff %.2e ; "
          "replicated rows agree across ranks: %s" % (dep, eps, err, not np.isnan(full).any()), flush=True)
    if not dep:
        for i in range(min(4, TUNE + DRAWS)):
            print("  transition %d: sharded depth %s size %s div %s maxdE %s | single depth %s size %s div %s maxdE %s" % (
                i, out["stats"][i, 0, :4], out["stats"][i, 2, :4], out["stats"][i, 3, :4], out["stats"][i, 7, :4],
                ref["stats"][i, 0, :4], ref["stats"][i, 2, :4], ref["stats"][i, 3, :4], ref["stats"][i, 7, :4]), flush=True)
    ok = same and dep and err < 1e-6 and eps < 1e-9
    print("SHARDED SAMPLER OK" if ok else "SHARDED SAMPLER MISMATCH", flush=True)
    e1.close()
eng.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
