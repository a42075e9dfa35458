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
            return json.load(open(p)), "measured"
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": wl["desc"], "sample": "%d chain-evals per step (one per host thread)" % T},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": T, "kind": "port",
                         "sample": "%d steps x %d chain-evals, full dataset per eval; C restatement of the "
                                   "reference's Theano op sequence (PyMC3 not installable)" % (args.steps, T)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def bench_gpu(args, wl):
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
    Q = synth.jittered_start(D, C, seed=1, model=model)
    gid = np.zeros(M, dtype=np.int32) if gw else d.gene_id

    if world == 1:
        eng = Engine(d.chi2, d.ld, None if gw else gid, G, d.c, C, model=model, device=local_rank)
        q_dev = eng.to_device_layout(Q)
        g_lo = 0
    else:
        from baghera_b200.sharding import shard_dataset
        sh = shard_dataset(d.chi2, d.ld, gid, G, rank, world)
        eng = Engine(sh.chi2, sh.ld, None if gw else sh.gene_id, sh.n_genes, d.c, C, model=model, device=local_rank, shard=sh.cfg)
        q_dev = eng.to_device_layout(Q if gw else sh.local_q(Q))
        lc = torch.tensor([eng.lik_const], dtype=torch.float64, device=dev)
        dist.all_reduce(lc)
        eng.set_lik_const(float(lc.item()))
        if args.exchange == "peer":
            from baghera_b200.sharding import connect_peers
            connect_peers(eng, dist, dev)
    Cp = eng.Cp
    logp = torch.empty(Cp, dtype=torch.float64, device=dev)
    grad = torch.zeros_like(q_dev)
    payload = torch.zeros((eng.payload_rows, Cp), dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step():
        if world == 1:
            eng.logp_grad(q_dev, logp, grad)
        elif args.exchange == "peer":
            # one launch per rank: local pass + one-shot all-reduce over NVLink peer memory + closing formulas
            eng.logp_grad_sharded(q_dev, logp, grad)
        else:
            eng.logp_grad_local(q_dev, grad, payload)
            dist.all_reduce(payload)
            eng.logp_grad_finish(q_dev, payload, logp, grad)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, L2 flushed between steps, per-step CUDA events ---------------
    K = args.steps
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    launches0 = eng.launch_count
    clocks = ClockSampler(local_rank).start() if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    wall0 = time.perf_counter()
    # a step is ONE kernel launch: the per-step events on the launch stream time exactly that launch (no
    # other event records inside the region - each record costs a few microseconds of stream time)
    for i in range(K):
        if not args.no_flush:
            flush.fill_(i & 0xFF)
        ev0[i].record()
        step()
        ev1[i].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - wall0
    clk = clocks.stop() if clocks else None
    launches = eng.launch_count - launches0
    step_ms = sum(e0.elapsed_time(e1) for e0, e1 in zip(ev0, ev1)) / K
    t = torch.tensor([step_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms = float(t.item())
    value = C / (step_ms * 1e-3)

    # ---- steady state (no flush, back-to-back: the sampler's regime, SNP stream L2 resident) -----
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_ss = min(max(K, 50), 2000)
    e0.record()
    for _ in range(n_ss):
        step()
    e1.record()
    torch.cuda.synchronize()
    ss_ms = e0.elapsed_time(e1) / n_ss

    line = None
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        lik_ms = step_ms        # one launch per step (with --exchange nccl: local pass + NCCL all-reduce + finish)
        Ml, Dl = eng.M, eng.D
        b_alg = (8.0 if gw else 12.0) * Ml + 16.0 * C * Dl + 8.0 * C   # SNP stream (no gene id in the gw model) + q + grad + logp
        achieved = b_alg / (lik_ms * 1e-3) / 1e9 if lik_ms > 0 else 0.0
        dp_instr = 5.0 * Ml * C                               # 5 DFMA per (SNP, chain): bgh_lik_device.cuh
        try:
            dfma_peak = eng.dfma_peak()
        except Exception:
            dfma_peak = None
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": achieved / peaks["hbm_gbs"], "traffic": ncu_traffic(args.workload, world), "peak_source": peak_kind,
                    "kernel": "logp_grad_fused_kernel (likelihood pass | grid barrier | finalize, one launch)" if world == 1
                    else "logp_grad_sharded_kernel (likelihood | finalize | peer all-reduce | finish, one launch)",
                    "kernel_ms": lik_ms,
                    "algorithmic_bytes_per_launch": b_alg,
                    "bytes_streamed_per_launch": 28.0 * Ml + 16.0 * C * Dl + 8.0 * C,
                    "note": "algorithmic bytes per SURVEY 8d (fp32 chi2/ld + int32 gene id = 12 B/SNP); the kernel streams a prepared "
                            "fp64 SoA (28 B/SNP) to keep conversions off the FP64 pipe. At C=64 the kernel is FP64-pipe bound "
                            "(5 DFMA per SNP x chain), not HBM bound: see fp64_roofline for the binding roof"}
        fp64 = {"achieved_ginstr_s": dp_instr / (lik_ms * 1e-3) / 1e9 if lik_ms > 0 else 0.0,
                "peak_ginstr_s": dfma_peak, "unit": "G FP64 lane-instr/s",
                "frac": (dp_instr / (lik_ms * 1e-3) / 1e9 / dfma_peak) if (dfma_peak and lik_ms > 0) else None,
                "peak_source": "bgh_dfma_peak microbenchmark, same process", "dp_instr_per_launch": dp_instr}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": wl["desc"], "M": M, "G": G, "chains": C,
                       "parallelism": "single GPU" if world == 1 else "gene-block SNP sharding x%d + %s of the %d x %d fp64 payload" % (
                           world, "in-kernel one-shot all-reduce over NVLink peer memory" if args.exchange == "peer" else "NCCL all-reduce", eng.payload_rows, Cp),
                       "l2": "not flushed" if args.no_flush else "flushed between steps (256 MiB write, outside the per-step events)",
                       "timing": "per-step CUDA events on the launch stream, max over ranks"},
            "batched_evals_per_sec": 1e3 / step_ms,
            "steady_state": {"ms_per_step": ss_ms, "value": C / (ss_ms * 1e-3), "note": "no L2 flush, back-to-back launches (sampler regime)"},
            "roofline": roofline, "fp64_roofline": fp64,
            "gpu_launches": int(launches), "clocks": clk, "wall_s_timed_region": wall,
        }
    # ---- e2e: host buffers through the C ABI (H2D + transposes + kernels + D2H every step) --------
    # N>1: every rank moves ITS rows of q / grad (3 shared + its gene block) over its own PCIe link and the
    # payload exchange runs inside the kernel; barrier on both sides, max over ranks.
    e2e_ok = world == 1 or args.exchange == "peer"
    if e2e_ok:
        Dl = eng.D
        Ql = Q if (world == 1 or gw) else sh.local_q(Q)
        n_e2e = int(min(max(K // 10, 8), 128))
        qh = torch.empty((n_e2e, C, Dl), dtype=torch.float64).pin_memory()
        gh = torch.empty((n_e2e, C, Dl), dtype=torch.float64).pin_memory()
        lh = torch.empty((n_e2e, C), dtype=torch.float64).pin_memory()
        qh.copy_(torch.from_numpy(np.broadcast_to(Ql, (n_e2e, C, Dl)).copy()))
        # warm-up: the whole buffer set once (allocates the staging; the first DMA to a fresh pinned page is slow in this VM)
        eng.logp_grad_host_ptr(n_e2e, qh.data_ptr(), lh.data_ptr(), gh.data_ptr())
        l0 = eng.launch_count
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        eng.logp_grad_host_ptr(n_e2e, qh.data_ptr(), lh.data_ptr(), gh.data_ptr())
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # check the e2e result against the device path
        lp_dev = logp[:C].cpu().numpy()
        assert np.allclose(lh[-1].numpy(), lp_dev, rtol=1e-12), "e2e result differs from device path"
        assert np.array_equal(gh[-1].numpy(), eng.from_device_layout(grad).cpu().numpy()), "e2e gradient differs from device path"
        if rank == 0:
            line["e2e"] = {"value": n_e2e * C / dt, "unit": UNIT, "h2d_bytes_per_step": C * Dl * 8,
                           "d2h_bytes_per_step": C * Dl * 8 + C * 8, "steps": n_e2e, "ms_per_step": dt / n_e2e * 1e3,
                           "api": "bgh_logp_grad_host_stream (pinned host q[C][D] -> logp[C], grad[C][D]; copies, layout "
                                  "transposes and kernels of consecutive steps overlapped on 3 streams, 4 evaluations in flight)",
                           "gpu_launches": eng.launch_count - l0}
            if world > 1:
                line["e2e"]["note"] = ("bytes are per rank: each rank copies its own rows of q/grad (3 shared + %d local genes) "
                                       "over its own PCIe link; wall clock, barrier before, max over ranks" % (Dl - 3))
    elif rank == 0:
        line["e2e"] = None
    if rank == 0 and world == 1 and not args.no_sampler:
        line["sampler"] = run_sampler_section(eng, args, C)
    if world > 1 and not gw and args.exchange == "peer" and not args.no_sampler:
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
            line["sharded_sampler"] = {"transitions": n_t + n_d, "batched_leapfrogs": nl, "seconds": float(t.item()),
                                       "us_per_batched_leapfrog": float(t.item()) / max(nl, 1) * 1e6,
                                       "mean_tree_accept": float(so["stats"][:, 1].mean()),
                                       "divergences_after_first_10": int(so["stats"][10:, 3].sum()),
                                       "note": "bgh_nuts_transition on sharded contexts: all %d chains advance in lock step, one "
                                               "sharded evaluation + 3 in-kernel peer exchanges per leapfrog; tools/"
                                               "sharded_sampler_check.py checks it against the single-GPU chains" % C}
        del smp
    if world > 1 and not gw:
        # Chains are independent units: every rank evaluates its OWN C chains on the FULL dataset, no collective
        # (weak scaling in chains: N x C chains in total).  Reported beside the sharded strong-scaling value.
        eng.close()
        eng = Engine(d.chi2, d.ld, d.gene_id, G, d.c, C, device=local_rank)
        q_full = eng.to_device_layout(Q)
        g_full = torch.zeros_like(q_full)
        for _ in range(max(args.warmup, 3)):
            eng.logp_grad(q_full, logp, g_full)
        n_rep = min(K, 500)
        r0 = [torch.cuda.Event(enable_timing=True) for _ in range(n_rep)]
        r1 = [torch.cuda.Event(enable_timing=True) for _ in range(n_rep)]
        torch.cuda.synchronize()
        dist.barrier()
        for i in range(n_rep):
            if not args.no_flush:
                flush.fill_(i & 0xFF)
            r0[i].record()
            eng.logp_grad(q_full, logp, g_full)
            r1[i].record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in zip(r0, r1)) / n_rep], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            line["chain_replicas"] = {"value": world * C / (float(t.item()) * 1e-3), "unit": UNIT, "scaling": "weak",
                                      "chains_total": world * C, "ms_per_step": float(t.item()), "steps": n_rep,
                                      "note": "every GPU evaluates its own %d chains on the full dataset (no collective); "
                                              "per-step CUDA events, L2 flushed, max over ranks" % C}
    if world > 1 and not args.no_sampler:
        # ESS/s at N GPUs: independent chain replicas (every rank runs its own C chains on the FULL dataset,
        # no collective; the gene-block sharded tick kernel is not built yet)
        sm = run_sampler_section(eng, args, C, seed=1 + rank)
        vec = torch.tensor([sm["ess_min_scalar"], sm["seconds_draws"], sm["ess_median_beta"],
                            sm["chain_leapfrogs_per_sec_post_tune"]], dtype=torch.float64, device=dev)
        allv = [torch.empty_like(vec) for _ in range(world)]
        dist.all_gather(allv, vec)
        if rank == 0:
            A = torch.stack(allv).cpu().numpy()
            secs = float(A[:, 1].max())
            sm.update({"parallelism": "%d independent replicas x %d chains (one per GPU, different seeds)" % (world, C),
                       "ess_min_scalar": float(A[:, 0].sum()), "ess_per_sec": float(A[:, 0].sum()) / secs,
                       "ess_median_beta": float(A[:, 2].sum()), "ess_median_beta_per_sec": float(A[:, 2].sum()) / secs,
                       "chain_leapfrogs_per_sec_post_tune": float(A[:, 3].sum()), "seconds_draws": secs})
            line["sampler"] = sm
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = run_cpu_baseline(d, C, target_seconds=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(line))
    eng.close()
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
