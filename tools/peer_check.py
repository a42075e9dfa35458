"""torchrun --nproc-per-node N tools/peer_check.py : sharded evaluation with the fused peer-memory exchange
against (a) the NCCL path and (b) the oracle on rank 0; then timing of both exchanges."""
import os, sys, time
import numpy as np, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baghera_b200 import synth
from baghera_b200.engine import Engine
from baghera_b200.sharding import shard_dataset, connect_peers, assemble_global_grad

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
M, G, C = int(os.environ.get("M", 1_100_000)), int(os.environ.get("G", 15_000)), int(os.environ.get("C", 64))
d = synth.make_arrays(M, G)
Q = synth.jittered_start(G + 3, C, seed=1)
sh = shard_dataset(d.chi2, d.ld, d.gene_id, G, rank, world)
eng = Engine(sh.chi2, sh.ld, sh.gene_id, sh.n_genes, d.c, C, device=lr, shard=sh.cfg)
lc = torch.tensor([eng.lik_const], dtype=torch.float64, device=dev)
dist.all_reduce(lc)
eng.set_lik_const(float(lc.item()))
connect_peers(eng, dist, dev)
q = eng.to_device_layout(sh.local_q(Q))
Cp = eng.Cp
lp1 = torch.empty(Cp, dtype=torch.float64, device=dev); g1 = torch.zeros_like(q)
lp2 = torch.empty(Cp, dtype=torch.float64, device=dev); g2 = torch.zeros_like(q)
payload = torch.zeros((eng.payload_rows, Cp), dtype=torch.float64, device=dev)

def nccl_step():
    eng.logp_grad_local(q, g1, payload)
    dist.all_reduce(payload)
    eng.logp_grad_finish(q, payload, lp1, g1)

def fused_step():
    eng.logp_grad_sharded(q, lp2, g2)

nccl_step(); fused_step()
torch.cuda.synchronize()
assert eng.peer_status() == 0, eng.peer_status()
a, b = lp1[:C].cpu().numpy(), lp2[:C].cpu().numpy()
ga, gb = g1.cpu().numpy(), g2.cpu().numpy()
err_lp = np.max(np.abs(a - b) / np.abs(a)); err_g = np.max(np.abs(ga - gb) / np.maximum(np.abs(ga), 1e-9 * np.abs(ga).max()))
# every rank must hold bit-identical logp
allp = [torch.empty_like(lp2) for _ in range(world)]
dist.all_gather(allp, lp2)
same = all(torch.equal(allp[0], t) for t in allp)
for name, fn in (("nccl", nccl_step), ("fused", fused_step)):
    for _ in range(20): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 500
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print("%s exchange: %.2f us/step (max over %d ranks)" % (name, float(t), world), flush=True)
assert eng.peer_status() == 0
# host entry on the sharded context: pinned local rows in, local gradient rows out, same bits as the device path
n_h = 6
qh = np.broadcast_to(sh.local_q(Q), (n_h, C, eng.D)).copy()
lph, gh = eng.logp_grad_host(qh)
fused_step(); torch.cuda.synchronize()
assert np.array_equal(lph[-1], lp2[:C].cpu().numpy()) and np.array_equal(lph[0], lph[-1])
gd = eng.from_device_layout(g2).cpu().numpy()
if not np.array_equal(gh[-1], gd):
    bad = np.argwhere(gh[-1] != gd)
    print("rank %d: host-entry gradient differs at %d of %d entries; rows %s; first: host %r device %r; D_local %d"
          % (rank, len(bad), gd.size, sorted(set(bad[:, 1].tolist()))[:12], gh[-1][tuple(bad[0])], gd[tuple(bad[0])], eng.D), flush=True)
    for k in range(n_h):
        print("   eval %d: mismatching entries %d" % (k, int((gh[k] != gd).sum())), flush=True)
assert np.array_equal(gh[-1], gd)
assert eng.peer_status() == 0
# gather the ranks' gradients on rank 0 and compare with the oracle (full dataset, first chains)
gl = [None] * world
dist.all_gather_object(gl, (gd, sh.row_coords()))
if rank == 0:
    from oracle import model_oracle as mo
    n_ref = min(C, 2)
    lp_ref, g_ref = mo.batch_closed(Q[:n_ref], d.chi2.astype(np.float64), d.ld.astype(np.float64), d.gene_id.astype(np.int64), d.c)
    full = np.full((C, G + 3), np.nan)
    for g_, rc in gl:
        full[:, rc] = g_
    e_lp = np.max(np.abs(lp2[:n_ref].cpu().numpy() - lp_ref) / np.abs(lp_ref))
    e_g = max(np.max(np.abs(full[i] - g_ref[i]) / np.maximum(np.abs(g_ref[i]), 1e-6 * np.abs(g_ref[i]).max())) for i in range(n_ref))
    print("fused exchange vs oracle: logp rel %.2e grad rel %.2e" % (e_lp, e_g), flush=True)
    assert e_lp < 1e-6 and e_g < 1e-6
    print("sharded host entry == device path", flush=True)
    print("fused vs nccl: logp rel %.2e grad rel %.2e ; logp bit-identical on all ranks: %s" % (err_lp, err_g, same), flush=True)
    assert err_lp < 1e-12 and err_g < 1e-9 and same
eng.close()
dist.destroy_process_group()
