#!/usr/bin/env python
"""bench.py — headline benchmark of the BAGHERA hot path on B200 (contract: see DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU restatement of the reference path

Metric: single-chain full-dataset logp+grad evaluations per second ("chain-evals/s") on the
UKBB-shaped synthetic workload BASELINE.json quotes: 1.1M SNPs x 15k genes x 64 chains (cfg3).
A step = one batched evaluation (all 64 chains, all SNPs) = 64 chain-evals.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "logp_grad_chain_evals_per_sec"
UNIT = "chain-evals/s"
WORKLOADS = {
    "cfg3": dict(M=1_100_000, G=15_000, C=64, desc="UKBB-shaped synthetic 1.1M SNPs x 15k genes, 64 chains, normal model"),
    "cfg2": dict(M=1_100_000, G=15_000, C=4, desc="UKBB-shaped synthetic 1.1M SNPs x 15k genes, 4 chains, normal model"),
    "cfg1": dict(M=20_000, G=200, C=4, desc="synthetic 20k SNPs x 200 genes, 4 chains, normal model"),
    "cfg5": dict(M=10_000_000, G=1, C=64, model="gw",
                 desc="genome-wide model (single h2 + intercept) on a 10M imputed-SNP synthetic set, 64 chains"),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            if float(d.get("hbm_gbs", 0)) > 0:
                return d, "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples SM clocks / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index=0, period=0.02):
        self.index, self.period = index, period
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._t = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            return self
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def ncu_traffic(workload, world):
    """dram bytes (read + write) per launch of the dominant kernel from the committed ncu --set full capture of
    this workload (profiles/ncu_traffic.json); None when no capture exists for the configuration."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            e = json.load(f).get("%s:%d" % (workload, world))
        return float(e["dram_read_bytes"] + e["dram_write_bytes"]) if e else None
    except (OSError, ValueError, KeyError):
        return None


def run_cpu_baseline(d, C, target_seconds=10.0, n_threads=0):
    """Times the C restatement of the reference CPU path (oracle/oracle_c.c) on the host cores."""
    from oracle.c_oracle import COracle, num_threads
    from baghera_b200 import synth
    T = n_threads or num_threads()
    co = COracle(d.chi2, d.ld, d.gene_id, d.G, d.c)
    Q = synth.jittered_start(d.G + 3, max(T, 1), seed=3)
    t0 = time.perf_counter()
    co.logp_grad_batch(Q, n_threads=T)
    t1 = time.perf_counter() - t0
    reps = int(max(1, min(5000, target_seconds / max(t1, 1e-4))))
    Qb = np.tile(Q, (reps, 1))
    t0 = time.perf_counter()
    co.logp_grad_batch(Qb, n_threads=T)
    dt = time.perf_counter() - t0
    n = Qb.shape[0]
    return {"value": n / dt, "unit": UNIT, "cores": T, "kind": "port",
            "sample": "%d chain-evals of the full %dx%d dataset (one chain per thread, %d threads, %.1f s); "
                      "C restatement of the Theano op sequence, not PyMC3 itself" % (n, d.M, d.G, T, dt),
            "ms_per_chain_eval_per_core": dt / n * T * 1e3}


def run_sampler_section(eng, args, C, seed=1):
    """ESS/s of the batched NUTS driver on the same dataset (secondary metric of BASELINE.json)."""
    import torch
    from baghera_b200 import stats as st
    from baghera_b200 import synth
    from baghera_b200.sampler import NutsSampler
    smp = NutsSampler(eng, seed=seed)
    smp.init(synth.jittered_start(eng.D, C, seed=1 + seed))
    l0 = eng.launch_count
    out = smp.sample(args.sampler_draws, args.sampler_tune, trace_dtype=torch.float32)
    tr = out["trace"]
    tune = args.sampler_tune
    secs = out["seconds_draws"]
    scal = {"W": torch.exp(tr[:, 0, :].double()).cpu().numpy(), "e": tr[:, 1, :].double().cpu().numpy(),
            "mi": torch.sigmoid(tr[:, 2, :].double()).cpu().numpy()}
    ess = {k: st.effective_n(v) for k, v in scal.items()}
    rhat = {k: st.gelman_rubin(v) for k, v in scal.items()}
    rng = np.random.default_rng(0)
    genes = rng.choice(eng.D - 3, size=min(64, eng.D - 3), replace=False) + 3
    beta_ess = [st.effective_n(tr[:, int(g), :].double().cpu().numpy()) for g in genes]
    post = out["stats"][tune:]
    leap = out["leapfrogs"]
    return {
        "chains": C, "tune": tune, "draws": args.sampler_draws,
        "seconds_tune": out["seconds_tune"], "seconds_draws": secs,
        "ess_min_scalar": min(ess.values()), "ess_per_sec": min(ess.values()) / secs, "ess": ess, "rhat": rhat,
        "ess_median_beta": float(np.median(beta_ess)), "ess_median_beta_per_sec": float(np.median(beta_ess)) / secs,
        "leapfrogs_per_draw_post_tune": float(leap[tune:].mean()), "leapfrogs_total": int(leap.sum()),
        "chain_leapfrogs_per_sec_post_tune": float(leap[tune:].sum()) * C / secs,
        "mean_tree_accept": float(post[:, 1].mean()), "divergences": int(post[:, 3].sum()),
        "mean_depth": float(post[:, 0].mean()), "gpu_launches": eng.launch_count - l0,
        "estimator": "pymc3 3.6 effective_n (Geyer initial sequences), post-tune draws only",
        "note": "short run: %d tuning draws do not fully converge the 15k-dim mass matrix; ESS/s is indicative" % tune,
    }


def bench_reference(args, wl):
    """Reference arm: the reference's own CPU implementation of the path.  PyMC3 3.6/theano cannot be
    installed here, so this times oracle/oracle_c.c (kind 'port') with every host thread."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from baghera_b200 import synth
    d = synth.make_arrays(wl["M"], wl["G"])
    from oracle.c_oracle import COracle, num_threads
    T = num_threads()
    co = COracle(d.chi2, d.ld, d.gene_id, d.G, d.c)
    # a step = T chain-evals (one per host thread): a bounded sample of the 64-chain batched evaluation
    Q = synth.jittered_start(d.G + 3, T, seed=3)
    for _ in range(args.warmup):
        co.logp_grad_batch(Q, n_threads=T)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.logp_grad_batch(Q, n_threads=T)
    dt = time.perf_counter() - t0
    value = args.steps * T / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.multi_gpu == "sharded" else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": wl["desc"], "sample": "%d chain-evals per step (one per host thread)" % T},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": T, "kind": "port",
                         "sample": "%d steps x %d chain-evals, full dataset per eval; C restatement of the "
                                   "reference's Theano op sequence (PyMC3 not installable)" % (args.steps, T)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def time_steps(step, K, warmup, flush, dist, world, dev, clocks=None):
    """W warm-up steps, then exactly K timed steps: per-step CUDA events on the launch stream (a step is one
    kernel launch; nothing else is recorded inside the region), L2 flushed between steps outside the events,
    barrier + synchronize on both sides, max over ranks.  Returns (ms per step, steady-state ms, wall s, clocks)."""
    import torch
    for _ in range(max(warmup, 3)):
        step()
    torch.cuda.synchronize()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    if clocks is not None:
        clocks.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    for i in range(K):
        if flush is not None:
            flush.fill_(i & 0xFF)
        ev0[i].record()
        step()
        ev1[i].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop() if clocks is not None else None
    step_ms = sum(e0.elapsed_time(e1) for e0, e1 in zip(ev0, ev1)) / K
    # steady state: no flush, back-to-back (the sampler's regime, SNP stream L2 resident)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_ss = min(max(K, 50), 2000)
    e0.record()
    for _ in range(n_ss):
        step()
    e1.record()
    torch.cuda.synchronize()
    ss_ms = e0.elapsed_time(e1) / n_ss
    t = torch.tensor([step_ms, ss_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0].item()), float(t[1].item()), wall, clk


def time_e2e(eng, Ql, C, K, warm_ref, dist, world, rank, dev):
    """Host buffers through the C ABI (bgh_logp_grad_host_stream): pinned q[n][C][D] in, logp / grad out, every
    step pays its H2D and D2H copies.  Barrier before, wall clock, max over ranks.  warm_ref = (logp, grad) of
    the device path for the same q: the host entry must reproduce them bit for bit."""
    import torch
    Dl = eng.D
    n = int(min(max(K // 10, 8), 128))
    qh = torch.empty((n, C, Dl), dtype=torch.float64).pin_memory()
    gh = torch.empty((n, C, Dl), dtype=torch.float64).pin_memory()
    lh = torch.empty((n, C), dtype=torch.float64).pin_memory()
    qh.copy_(torch.from_numpy(np.broadcast_to(Ql, (n, C, Dl)).copy()))
    # warm-up: the whole buffer set once (allocates the staging; the first DMA to a fresh pinned page is slow in this VM)
    eng.logp_grad_host_ptr(n, qh.data_ptr(), lh.data_ptr(), gh.data_ptr())
    l0 = eng.launch_count
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    eng.logp_grad_host_ptr(n, qh.data_ptr(), lh.data_ptr(), gh.data_ptr())
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    logp, grad = warm_ref
    assert np.allclose(lh[-1].numpy(), logp[:C].cpu().numpy(), rtol=1e-12), "e2e result differs from device path"
    assert np.array_equal(gh[-1].numpy(), eng.from_device_layout(grad).cpu().numpy()), "e2e gradient differs from device path"
    return {"seconds": dt, "steps": n, "ms_per_step": dt / n * 1e3, "h2d_bytes_per_step": C * Dl * 8,
            "d2h_bytes_per_step": C * Dl * 8 + C * 8, "gpu_launches": eng.launch_count - l0,
            "api": "bgh_logp_grad_host_stream (pinned host q[C][D] -> logp[C], grad[C][D]; copies, layout transposes and "
                   "kernels of consecutive steps overlapped, copies alternate between two streams per direction, 4 groups in flight)"}


def bench_gpu(args, wl):
    """N = 1: one GPU evaluates the 64 chains.  N > 1 (one process per GPU): the headline is CHAIN REPLICAS - chains
    are the independent units of this path, every GPU evaluates its own 64 chains over the full dataset with no
    data-path collective (weak scaling, N x 64 chains) - and the gene-block SNP sharding of BASELINE.json's config 4
    (64 chains in total, one in-kernel peer-memory all-reduce per evaluation, strong scaling) is measured in the same
    run and reported under "sharded" (--multi-gpu sharded makes it the headline instead)."""
    import torch
    import torch.distributed as dist

    from baghera_b200 import synth
    from baghera_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    M, G, C = wl["M"], wl["G"], wl["C"]
    model = wl.get("model", "gene")
    gw = model == "gw"
    if gw:
        args.no_sampler = True          # the ESS/s section and the CPU port are for the gene-level model
        args.no_cpu_baseline = True
    d = synth.make_arrays(M, G)
    D = 2 if gw else G + 3
    gid = np.zeros(M, dtype=np.int32) if gw else d.gene_id
    K = args.steps
    flush = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    peaks, peak_kind = measured_peaks()
    l2_note = "not flushed" if args.no_flush else "flushed between steps (256 MiB write, outside the per-step events)"

    def roofs(eng, step_ms, kernel, traffic):
        Ml, Dl = eng.M, eng.D
        b_alg = (8.0 if gw else 12.0) * Ml + 16.0 * C * Dl + 8.0 * C   # SNP stream (no gene id in the gw model) + q + grad + logp
        achieved = b_alg / (step_ms * 1e-3) / 1e9
        dp_instr = 5.0 * Ml * C                               # 5 DFMA per (SNP, chain): bgh_lik_device.cuh
        try:
            dfma_peak = eng.dfma_peak()
        except Exception:
            dfma_peak = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_kind,
                    "kernel": kernel, "kernel_ms": step_ms, "algorithmic_bytes_per_launch": b_alg,
                    "bytes_streamed_per_launch": 28.0 * Ml + 16.0 * C * Dl + 8.0 * C,
                    "note": "per GPU; algorithmic bytes per SURVEY 8d (fp32 chi2/ld + int32 gene id = 12 B/SNP); the kernel streams a "
                            "prepared fp64 SoA (28 B/SNP) to keep conversions off the FP64 pipe. At C=64 the kernel is FP64-pipe "
                            "bound (5 DFMA per SNP x chain), not HBM bound: see fp64_roofline for the binding roof"}
        fp64 = {"achieved_ginstr_s": dp_instr / (step_ms * 1e-3) / 1e9, "peak_ginstr_s": dfma_peak,
                "unit": "G FP64 lane-instr/s", "frac": (dp_instr / (step_ms * 1e-3) / 1e9 / dfma_peak) if dfma_peak else None,
                "peak_source": "bgh_dfma_peak microbenchmark, same process", "dp_instr_per_launch": dp_instr}
        return roofline, fp64

    line = None
    sharded_primary = world > 1 and args.multi_gpu == "sharded"
    if args.multi_gpu == "sharded":
        args.no_sharded = False

    # ------------------------------------------------------------------------------------------------------
    # (A) every GPU: its own C chains on the full dataset (N = 1: the single-GPU measurement)
    # ------------------------------------------------------------------------------------------------------
    Q = synth.jittered_start(D, C, seed=1 + rank, model=model)
    eng = Engine(d.chi2, d.ld, None if gw else gid, G, d.c, C, model=model, device=local_rank)
    q_dev = eng.to_device_layout(Q)
    logp = torch.empty(eng.Cp, dtype=torch.float64, device=dev)
    grad = torch.zeros_like(q_dev)

    def step_full():
        eng.logp_grad(q_dev, logp, grad)

    l0 = eng.launch_count
    step_ms, ss_ms, wall, clk = time_steps(step_full, K, args.warmup, flush, dist, world, dev,
                                           ClockSampler(local_rank) if rank == 0 else None)
    launches_full = K                                        # one fused kernel per timed step (eng.launch_count agrees)
    assert eng.launch_count - l0 >= K
    e2e_full = time_e2e(eng, Q, C, K, (logp, grad), dist, world, rank, dev)
    if rank == 0:
        roofline, fp64 = roofs(eng, step_ms, "logp_grad_fused_kernel (likelihood pass | grid barrier | finalize, one launch)",
                               ncu_traffic(args.workload, 1))
        line = {
            "metric": METRIC, "value": world * C / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if args.multi_gpu == "sharded" else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "M": M, "G": G, "chains": C, "chains_total": world * C,
                       "parallelism": "single GPU" if world == 1 else
                       "chain replicas: each of the %d GPUs evaluates its own %d chains over the full dataset, no data-path "
                       "collective (chains are independent units); the gene-block sharded path is under 'sharded'" % (world, C),
                       "l2": l2_note, "timing": "per-step CUDA events on the launch stream, barrier on both sides, max over ranks"},
            "batched_evals_per_sec": world * 1e3 / step_ms,
            "steady_state": {"ms_per_step": ss_ms, "value": world * C / (ss_ms * 1e-3),
                             "note": "no L2 flush, back-to-back launches (sampler regime)"},
            "roofline": roofline, "fp64_roofline": fp64, "gpu_launches": int(world * launches_full), "clocks": clk,
            "wall_s_timed_region": wall,
            "e2e": dict(e2e_full, value=world * e2e_full["steps"] * C / e2e_full["seconds"], unit=UNIT),
        }
        if world > 1:
            line["e2e"]["note"] = ("every rank moves its own q/grad (bytes are per rank); the GPUs of this box share its "
                                   "host<->device bandwidth, so the end-to-end figure does not grow with N")

    # ESS/s of the batched NUTS driver (N > 1: independent replicas of C chains per GPU, different seeds)
    if not args.no_sampler:
        sm = run_sampler_section(eng, args, C, seed=1 + rank)
        if world > 1:
            vec = torch.tensor([sm["ess_min_scalar"], sm["seconds_draws"], sm["ess_median_beta"],
                                sm["chain_leapfrogs_per_sec_post_tune"]], dtype=torch.float64, device=dev)
            allv = [torch.empty_like(vec) for _ in range(world)]
            dist.all_gather(allv, vec)
            A = torch.stack(allv).cpu().numpy()
            secs = float(A[:, 1].max())
            sm.update({"parallelism": "%d independent replicas x %d chains (one per GPU, different seeds)" % (world, C),
                       "ess_min_scalar": float(A[:, 0].sum()), "ess_per_sec": float(A[:, 0].sum()) / secs,
                       "ess_median_beta": float(A[:, 2].sum()), "ess_median_beta_per_sec": float(A[:, 2].sum()) / secs,
                       "chain_leapfrogs_per_sec_post_tune": float(A[:, 3].sum()), "seconds_draws": secs})
        if rank == 0:
            line["sampler"] = sm
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = run_cpu_baseline(d, C, target_seconds=args.cpu_seconds)
    eng.close()
    del eng, q_dev, grad

    # ------------------------------------------------------------------------------------------------------
    # (B) N > 1: gene-block SNP sharding (BASELINE.json config 4): 64 chains in total, strong scaling
    # ------------------------------------------------------------------------------------------------------
    if world > 1 and not args.no_sharded:
        from baghera_b200.sharding import connect_peers, shard_dataset
        Qs = synth.jittered_start(D, C, seed=1, model=model)          # the same chains on every rank
        sh = shard_dataset(d.chi2, d.ld, gid, G, rank, world)
        eng = Engine(sh.chi2, sh.ld, None if gw else sh.gene_id, sh.n_genes, d.c, C, model=model, device=local_rank, shard=sh.cfg)
        Ql = Qs if gw else sh.local_q(Qs)
        q_dev = eng.to_device_layout(Ql)
        lc = torch.tensor([eng.lik_const], dtype=torch.float64, device=dev)
        dist.all_reduce(lc)
        eng.set_lik_const(float(lc.item()))
        if args.exchange == "peer":
            connect_peers(eng, dist, dev)
        logp = torch.empty(eng.Cp, dtype=torch.float64, device=dev)
        grad = torch.zeros_like(q_dev)
        payload = torch.zeros((eng.payload_rows, eng.Cp), dtype=torch.float64, device=dev)

        def step_sharded():
            if args.exchange == "peer":
                # one launch per rank: local pass + one-shot all-reduce over NVLink peer memory + closing formulas
                eng.logp_grad_sharded(q_dev, logp, grad)
            else:
                eng.logp_grad_local(q_dev, grad, payload)
                dist.all_reduce(payload)
                eng.logp_grad_finish(q_dev, payload, logp, grad)

        Ks = min(K, 1000)
        s_ms, s_ss, s_wall, _ = time_steps(step_sharded, Ks, args.warmup, flush, dist, world, dev)
        s_e2e = time_e2e(eng, Ql, C, Ks, (logp, grad), dist, world, rank, dev) if args.exchange == "peer" else None
        shd = None
        if rank == 0:
            kern = ("logp_grad_sharded_kernel (likelihood | finalize | peer all-reduce | finish, one launch)" if args.exchange == "peer"
                    else "local pass + NCCL all-reduce + finish (3 launches)")
            r2, f2 = roofs(eng, s_ms, kern, None)
            shd = {"value": C / (s_ms * 1e-3), "unit": UNIT, "scaling": "strong", "chains_total": C, "steps": Ks,
                   "ms_per_step": s_ms, "steady_state": {"ms_per_step": s_ss, "value": C / (s_ss * 1e-3)},
                   "parallelism": "gene-block SNP sharding x%d + %s of the %d x %d fp64 payload" % (
                       world, "in-kernel one-shot all-reduce over NVLink peer memory" if args.exchange == "peer" else "NCCL all-reduce",
                       eng.payload_rows, eng.Cp),
                   "roofline": r2, "fp64_roofline": f2, "gpu_launches": Ks * (1 if args.exchange == "peer" else 2)}
            if s_e2e is not None:
                shd["e2e"] = dict(s_e2e, value=s_e2e["steps"] * C / s_e2e["seconds"], unit=UNIT,
                                  note="bytes are per rank: each rank copies its own rows of q/grad (3 shared + %d local genes)" % (eng.D - 3))
        if not gw and args.exchange == "peer" and not args.no_sampler:
            # the sampler on the gene-block shards (lock-step driver: every rank owns its rows of q, per-chain sums
            # all-reduced over peer memory inside the sampler's scalar kernels; identical decisions on every rank)
            from baghera_b200.sampler import NutsSampler
            smp = NutsSampler(eng, seed=1, gene_lo=sh.gene_lo, n_dim_total=D)
            smp.init(sh.local_q(synth.jittered_start(D, C, seed=2)))
            dist.barrier()
            n_t, n_d = 20, 10
            t0 = time.perf_counter()
            so = smp.sample_lockstep(n_d, n_t)
            secs = time.perf_counter() - t0
            t = torch.tensor([secs], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            if rank == 0:
                nl = int(so["leapfrogs"].sum())
                shd["sampler"] = {"transitions": n_t + n_d, "batched_leapfrogs": nl, "seconds": float(t.item()),
                                  "us_per_batched_leapfrog": float(t.item()) / max(nl, 1) * 1e6,
                                  "mean_tree_accept": float(so["stats"][:, 1].mean()),
                                  "note": "bgh_nuts_transition on sharded contexts: all %d chains advance in lock step, one sharded "
                                          "evaluation + in-kernel peer exchanges of the per-chain sums per leapfrog; "
                                          "tools/sharded_sampler_check.py checks it against the single-GPU chains" % C}
            del smp
        if rank == 0:
            if sharded_primary:
                # --multi-gpu sharded: BASELINE.json's config 4 as the headline, replicas beside it
                rep = {k: line[k] for k in ("value", "ms_per_step", "steady_state", "e2e", "batched_evals_per_sec")}
                rep.update(scaling="weak", chains_total=world * C, parallelism=line["config"]["parallelism"])
                line.update({"value": shd["value"], "ms_per_step": shd["ms_per_step"], "scaling": "strong", "steps": Ks,
                             "steady_state": shd["steady_state"], "roofline": shd["roofline"], "fp64_roofline": shd["fp64_roofline"],
                             "batched_evals_per_sec": 1e3 / shd["ms_per_step"], "gpu_launches": shd["gpu_launches"],
                             "e2e": shd.get("e2e"), "chain_replicas": rep})
                line["config"].update(parallelism=shd["parallelism"], chains_total=C)
            else:
                line["sharded"] = shd
        eng.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: payload all-reduce fused into the kernel over NVLink peer memory (default) or a separate NCCL call")
    ap.add_argument("--multi-gpu", default="replicas", choices=["replicas", "sharded"],
                    help="N>1 headline: chain replicas (weak scaling, default) or the gene-block sharded path (strong scaling); "
                         "both are measured and reported")
    ap.add_argument("--no-sharded", action="store_true", help="N>1: skip the gene-block sharded measurements")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-sampler", action="store_true", help="skip the ESS/s section")
    ap.add_argument("--sampler-tune", type=int, default=300)
    ap.add_argument("--sampler-draws", type=int, default=200)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if args.steps > 5000:
            args.steps = 5000        # a step is ~7 ms (one chain-eval per host thread in parallel); keep the run bounded
        bench_reference(args, wl)
    else:
        bench_gpu(args, wl)


if __name__ == "__main__":
    main()
