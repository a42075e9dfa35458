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


def run_sampler_section(eng, args, C, seed=1, shard=None, D_total=None, dist=None, dev=None):
    """ESS/s of the batched NUTS driver on the same dataset (secondary metric of BASELINE.json).  Sharded engines: the
    persistent kernel on the gene-block shards (64 chains in total); the scalar parameters are replicated rows, so
    rank 0 holds their trace; the per-gene ESS is taken over the rank's own genes."""
    import torch
    from baghera_b200 import stats as st
    from baghera_b200 import synth
    from baghera_b200.sampler import NutsSampler
    world = dist.get_world_size() if (dist is not None and shard is not None) else 1
    if shard is not None:
        smp = NutsSampler(eng, seed=seed, row_coords=shard.row_coords(), n_dim_total=D_total)
        smp.init(shard.local_q(synth.jittered_start(D_total, C, seed=1 + seed)))
        dist.barrier()
    else:
        smp = NutsSampler(eng, seed=seed)
        smp.init(synth.jittered_start(eng.D, C, seed=1 + seed))
    l0 = eng.launch_count
    out = smp.sample(args.sampler_draws, args.sampler_tune, trace_dtype=torch.float32)
    tr = out["trace"]
    tune = args.sampler_tune
    secs, secs_total = out["seconds_draws"], out["seconds_total"]
    if world > 1:
        t = torch.tensor([secs, secs_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        secs, secs_total = float(t[0].item()), float(t[1].item())
    scal = {"W": torch.exp(tr[:, 0, :].double()).cpu().numpy(), "e": tr[:, 1, :].double().cpu().numpy(),
            "mi": torch.sigmoid(tr[:, 2, :].double()).cpu().numpy()}
    ess = {k: st.effective_n(v) for k, v in scal.items()}
    rhat = {k: st.gelman_rubin(v) for k, v in scal.items()}
    rng = np.random.default_rng(0)
    first = 3 + (shard.n_replicated if shard is not None else 0)
    genes = rng.choice(eng.D - first, size=min(64, eng.D - first), replace=False) + first
    beta_ess = [st.effective_n(tr[:, int(g), :].double().cpu().numpy()) for g in genes]
    post = out["stats"][tune:]
    leap = out["leapfrogs"]
    ticks = int(out.get("ticks", 0))
    return {
        "chains": C, "tune": tune, "draws": args.sampler_draws, "ticks": ticks,
        "us_per_tick": secs_total / max(ticks, 1) * 1e6,
        "engine": "fused persistent kernel, gene-block sharded x%d (bgh_tick2.cu)" % world if world > 1 else "bgh_nuts_run (persistent kernel)",
        "seconds_total": secs_total, "seconds_tune": out["seconds_tune"], "seconds_draws": secs,
        "seconds_note": "seconds_total: max over ranks; " + out.get("seconds_note", ""),
        "ess_min_scalar": min(ess.values()), "ess_per_sec": min(ess.values()) / secs, "ess": ess, "rhat": rhat,
        "ess_median_beta": float(np.median(beta_ess)), "ess_median_beta_per_sec": float(np.median(beta_ess)) / secs,
        "leapfrogs_per_draw_post_tune": float(leap[tune:].mean()), "leapfrogs_total": int(leap.sum()),
        "chain_leapfrogs_per_sec": float(out["stats"][:, 2].sum()) / secs_total,
        "chain_leapfrogs_per_sec_post_tune": float(leap[tune:].sum()) * C / secs,
        "mean_tree_accept": float(post[:, 1].mean()), "divergences": int(post[:, 3].sum()),
        "mean_depth": float(post[:, 0].mean()), "gpu_launches": eng.launch_count - l0,
        "estimator": "pymc3 3.6 effective_n (Geyer initial sequences), post-tune draws only",
        "note": "short run: %d tuning draws do not fully converge the 15k-dim mass matrix; ESS/s is indicative (the runs at "
                "BASELINE.json's sweeps/burn-in are recorded in profiles/)" % tune,
    }


def bench_config(wl, C):
    """The `config` object, identical in both arms (the driver compares them)."""
    return {"workload": wl["desc"], "M": wl["M"], "G": wl["G"], "chains": C, "chains_total": C}


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
    C = wl["C"]
    co = COracle(d.chi2, d.ld, d.gene_id, d.G, d.c)
    # a step = one batched evaluation = C chain-evals of the full dataset, one chain per host thread at a time
    Q = synth.jittered_start(d.G + 3, C, seed=3)
    for _ in range(args.warmup):
        co.logp_grad_batch(Q, n_threads=T)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        co.logp_grad_batch(Q, n_threads=T)
    dt = time.perf_counter() - t0
    value = args.steps * C / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": bench_config(wl, C),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": T, "kind": "port",
                         "sample": "%d steps x %d chain-evals (one batched evaluation of %d chains on %d host threads), full dataset "
                                   "per eval; C restatement of the reference's Theano op sequence (PyMC3 not installable)" % (
                                       args.steps, C, C, T)},
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
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()          # ranks wait for each other inside the sharded kernel: start the loop together
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


def pcie_floor_ms(nbytes_each_way, dev, reps=32):
    """Both-direction copy floor of this box for one step's transfers: nbytes H2D and nbytes D2H of pinned memory on two
    streams, concurrently, back to back (tools/pcie_probe.py inside the bench)."""
    import torch
    n = max(int(nbytes_each_way) // 8, 1)
    h_in = torch.empty((4, n), dtype=torch.float64).pin_memory()
    h_out = torch.empty((4, n), dtype=torch.float64).pin_memory()
    h_in.fill_(1.0); h_out.fill_(0.0)
    d_in = torch.empty(n, dtype=torch.float64, device=dev)
    d_out = torch.zeros(n, dtype=torch.float64, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    best = None
    for rep in range(3):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for i in range(reps):
            with torch.cuda.stream(s1):
                d_in.copy_(h_in[i % 4], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[i % 4].copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(dev)
        dt = (time.perf_counter() - t0) / reps * 1e3
        best = dt if best is None else min(best, dt)
    return best


def time_e2e(eng, Ql, C, warm_ref, dist, world, rank, dev, n=64):
    """Host buffers through the C ABI (bgh_logp_grad_host_stream): pinned q[n][C][D] in, logp / grad out, every
    step pays its H2D and D2H copies.  n >= 64 evaluations whatever --steps is.  Barrier before, wall clock, max over
    ranks.  warm_ref = (logp, grad) of the device path for the same q: the host entry must reproduce them bit for bit."""
    import torch
    Dl = eng.D
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
    floor = pcie_floor_ms(C * Dl * 8, dev)
    t = torch.tensor([dt, floor], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, floor = float(t[0].item()), float(t[1].item())
    logp, grad = warm_ref
    assert np.allclose(lh[-1].numpy(), logp[:C].cpu().numpy(), rtol=1e-12), "e2e result differs from device path"
    assert np.array_equal(gh[-1].numpy(), eng.from_device_layout(grad).cpu().numpy()), "e2e gradient differs from device path"
    return {"seconds": dt, "steps": n, "ms_per_step": dt / n * 1e3, "h2d_bytes_per_step": C * Dl * 8,
            "d2h_bytes_per_step": C * Dl * 8 + C * 8, "gpu_launches": eng.launch_count - l0,
            "pcie_floor_ms": floor, "frac_of_pcie_floor": floor / (dt / n * 1e3),
            "pcie_floor_note": "one step's bytes copied both ways concurrently (pinned, two streams, back to back) on this box%s"
                               % (", all ranks at the same time (max over ranks)" if world > 1 else ""),
            "api": "bgh_logp_grad_host_stream (pinned host q[C][D] -> logp[C], grad[C][D]; copies, layout transposes and "
                   "kernels of consecutive steps overlapped, copies alternate between two streams per direction, 4 groups in flight)"}


def oracle_parity(Q, lp_dev, grad_cd, d, gw, n_ref=2):
    """logp / gradient of the first chains against the numpy oracle (oracle/model_oracle.py) on the full dataset."""
    from oracle import model_oracle as mo
    chi2, ld = d.chi2.astype(np.float64), d.ld.astype(np.float64)
    e_lp = e_g = 0.0
    for i in range(n_ref):
        if gw:
            lp_ref, g_ref = mo.gw_logp_grad_closed(Q[i], chi2, ld, d.c)
        else:
            lp_ref, g_ref = mo.logp_grad_closed(Q[i], chi2, ld, d.gene_id.astype(np.int64), d.c)
        e_lp = max(e_lp, abs(lp_dev[i] - lp_ref) / abs(lp_ref))
        floor = 1e-6 * np.max(np.abs(g_ref))
        e_g = max(e_g, float(np.max(np.abs(grad_cd[i] - g_ref) / np.maximum(np.abs(g_ref), floor))))
    return {"logp_rel": float(e_lp), "grad_rel": float(e_g), "chains_checked": n_ref, "tolerance": 1e-6,
            "against": "oracle/model_oracle.py closed form on the full dataset (CPU, fp64); the oracle itself is pinned to its own "
                       "independent forms, scipy.stats and finite differences only: PyMC3 3.6 is not installable (parity: oracle-only)"}


def bench_gpu(args, wl):
    """N = 1: one GPU evaluates the 64 chains over the full dataset.  N > 1 (one process per GPU): BASELINE.json's config 4,
    gene-block SNP sharding of the SAME 64 chains (strong scaling): every rank streams its gene block, the per-chain sums
    cross NVLink in one LL-protocol peer exchange fused into the kernel, per-gene gradients stay local.  Chain replicas
    (every GPU its own 64 chains, no collective) are a product feature (--devices) and are reported under
    "chain_replicas" with --with-replicas only."""
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
        # register-file ceiling of the built kernel's hot loop (a DFMA that reads three fresh operands takes 3 pipe
        # cycles instead of 2): static analysis of the SASS of the library that just ran
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            from sass_dfma_reads import kernel_ceiling
            cl, cpl = (eng.Cp, 1) if eng.Cp <= 32 else (32, 2)
            gam = 1 if model == "gamma" else 0
            mangled = ("_ZN3bgh23logp_grad_chainx_kernelILi%dELi%dELb%dEEEvNS_11Eval2ParamsE" if world > 1 else
                       "_ZN3bgh22logp_grad_fused_kernelILi%dELi%dELb%dEEEvNS_11FusedParamsE") % (cl, cpl, gam)
            fp64["rf_ceiling"] = kernel_ceiling(os.path.join(ROOT, "baghera_b200", "libbaghera_b200.so"), mangled)
        except Exception:
            fp64["rf_ceiling"] = None
        return roofline, fp64

    Q = synth.jittered_start(D, C, seed=1, model=model)           # the same chains at every N
    shard = None
    if world == 1:
        eng = Engine(d.chi2, d.ld, None if gw else d.gene_id, G, d.c, C, model=model, device=local_rank)
        Ql = Q
        kernel = "logp_grad_fused_kernel (likelihood pass | grid barrier | finalize, one launch)"
        parallelism = "single GPU"
    else:
        from baghera_b200.sharding import connect_peers, shard_dataset, shard_snps_only
        shard = shard_snps_only(d.chi2, d.ld, rank, world) if gw else shard_dataset(d.chi2, d.ld, d.gene_id, G, rank, world)
        eng = Engine(shard.chi2, shard.ld, None if gw else shard.gene_id, shard.n_genes, d.c, C, model=model, device=local_rank,
                     shard=shard.cfg)
        Ql = Q if gw else shard.local_q(Q)
        lc = torch.tensor([eng.lik_const], dtype=torch.float64, device=dev)
        dist.all_reduce(lc)
        eng.set_lik_const(float(lc.item()))
        connect_peers(eng, dist, dev)
        kernel = "logp_grad_chainx_kernel (likelihood pass | grid barrier | per-chain closing with the LL peer exchange, one launch)"
        parallelism = ("gene-block SNP sharding x%d of the same %d chains: whole genes per rank, %d replicated gene(s) (NonCoding) "
                       "split evenly; one in-kernel LL-protocol exchange of %d fp64 per chain over NVLink peer memory per evaluation"
                       % (world, C, shard.n_replicated, 4 + shard.n_replicated))
    q_dev = eng.to_device_layout(Ql)
    logp = torch.empty(eng.Cp, dtype=torch.float64, device=dev)
    grad = torch.zeros_like(q_dev)

    def step():
        if world == 1:
            eng.logp_grad(q_dev, logp, grad)
        else:
            eng.logp_grad_sharded(q_dev, logp, grad)

    l0 = eng.launch_count
    step_ms, ss_ms, wall, clk = time_steps(step, K, args.warmup, flush, dist, world, dev, ClockSampler(local_rank) if rank == 0 else None)
    assert eng.launch_count - l0 >= K
    if world > 1:
        assert eng.peer_status() == 0, "peer exchange timed out"
    # parity of THIS run's result (every N): first chains against the CPU oracle on the full dataset
    step()
    torch.cuda.synchronize()
    g_cd = eng.from_device_layout(grad).cpu().numpy()
    lp_host = logp[:C].cpu().numpy()
    parity = None
    if world > 1:
        gathered = [None] * world
        dist.all_gather_object(gathered, (g_cd[:2], None if gw else shard.row_coords(), lp_host[:2]))
        if rank == 0:
            full = np.full((2, D), np.nan)
            for g_, rc, lp_ in gathered:
                if gw:
                    full[:] = g_
                else:
                    full[:, rc] = g_
                assert np.array_equal(lp_, gathered[0][2]), "logp differs between ranks"
            parity = oracle_parity(Q, lp_host, full, d, gw)
    elif rank == 0:
        parity = oracle_parity(Q, lp_host, g_cd, d, gw)
    if parity is not None:
        assert parity["logp_rel"] < 1e-6 and parity["grad_rel"] < 1e-6, parity
    e2e = time_e2e(eng, Ql, C, (logp, grad), dist, world, rank, dev, n=args.e2e_evals)
    line = None
    if rank == 0:
        roofline, fp64 = roofs(eng, step_ms, kernel, ncu_traffic(args.workload, world))
        cfg = bench_config(wl, C)
        line = {
            "metric": METRIC, "value": C / (step_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
            "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "parallelism": parallelism, "l2": l2_note,
            "timing": "per-step CUDA events on the launch stream, barrier on both sides, max over ranks",
            "batched_evals_per_sec": 1e3 / step_ms,
            "steady_state": {"ms_per_step": ss_ms, "value": C / (ss_ms * 1e-3),
                             "note": "no L2 flush, back-to-back launches (sampler regime)"},
            "roofline": roofline, "fp64_roofline": fp64, "gpu_launches": int(K), "clocks": clk,
            "wall_s_timed_region": wall, "parity": parity,
            "e2e": dict(e2e, value=e2e["steps"] * C / e2e["seconds"], unit=UNIT),
        }
        if world > 1:
            line["e2e"]["note"] = ("bytes are per rank: every rank copies its own rows of q / grad (%d local rows of %d); the GPUs of "
                                   "one box share its host<->device bandwidth" % (eng.D, D))

    # ESS/s of the batched NUTS driver (N > 1: the fused persistent kernel on the gene-block shards, the same 64 chains)
    if not args.no_sampler:
        sm = run_sampler_section(eng, args, C, seed=1, shard=shard, D_total=D, dist=dist if world > 1 else None, dev=dev)
        if rank == 0:
            line["sampler"] = sm
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = run_cpu_baseline(d, C, target_seconds=args.cpu_seconds)
        if not args.no_sampler and not args.no_cfg1:
            line["cfg1_sampler"] = run_cfg1_sampler(args, local_rank)
            line["cpu_baseline"]["sampler"] = line["cfg1_sampler"]["cpu"]
    eng.close()
    del eng, q_dev, grad

    # chain replicas (no data-path collective): every GPU its own C chains over the full dataset
    if world > 1 and args.with_replicas:
        eng = Engine(d.chi2, d.ld, None if gw else d.gene_id, G, d.c, C, model=model, device=local_rank)
        Qr = synth.jittered_start(D, C, seed=1 + rank, model=model)
        q_dev = eng.to_device_layout(Qr)
        logp = torch.empty(eng.Cp, dtype=torch.float64, device=dev)
        grad = torch.zeros_like(q_dev)
        r_ms, r_ss, _, _ = time_steps(lambda: eng.logp_grad(q_dev, logp, grad), min(K, 500), args.warmup, flush, dist, world, dev)
        if rank == 0:
            line["chain_replicas"] = {"value": world * C / (r_ms * 1e-3), "unit": UNIT, "scaling": "weak", "chains_total": world * C,
                                      "ms_per_step": r_ms, "steady_state_ms": r_ss,
                                      "parallelism": "each of the %d GPUs evaluates its own %d chains over the full dataset" % (world, C)}
        eng.close()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def _cfg1_cpu_chain(args):
    """One chain of the CPU oracle NUTS on cfg1 (worker of run_cfg1_sampler; one process per chain like pm.sample(cores=N))."""
    chain, draws, tune, seed = args
    from baghera_b200 import synth
    from oracle import nuts_oracle as no
    from oracle.c_oracle import COracle
    d = synth.make_arrays(20_000, 200)
    starts = synth.jittered_start(d.G + 3, 4, seed=31)
    co = COracle(d.chi2, d.ld, d.gene_id, d.G, d.c)
    ch = no.NutsChain(co.logp_grad, starts[chain], starts.mean(axis=0), np.random.default_rng([seed, chain]), target_accept=0.90)
    trace = np.empty((draws, d.G + 3))
    tree = np.empty(tune + draws)
    t0 = time.perf_counter()
    for i in range(tune + draws):
        if i == tune:
            ch.stop_tuning()
        st = ch.step()
        tree[i] = st["tree_size"]
        if i >= tune:
            trace[i - tune] = ch.q
    return trace, {"tree_size": tree}, time.perf_counter() - t0


def run_cfg1_sampler(args, device):
    """BASELINE.json config 1 exactly (synthetic 20k SNPs x 200 genes, 4 chains, sweeps 1000, burn-in 250): the CUDA sampler
    and the CPU oracle NUTS (oracle/nuts_oracle.py: PyMC3 3.6 semantics in numpy; one process per chain, --n-cores 4) on the
    same data; ESS/s of both and the z-scores of the posterior means (difference / combined MC standard error)."""
    import multiprocessing as mp

    import torch
    from baghera_b200 import stats as st
    from baghera_b200 import synth
    from baghera_b200.engine import Engine
    from baghera_b200.sampler import NutsSampler
    draws, tune, chains, n_cores = 1000, 250, 4, 4
    d = synth.make_arrays(20_000, 200)
    eng = Engine(d.chi2, d.ld, d.gene_id, d.G, d.c, chains, device=device)
    smp = NutsSampler(eng, seed=17)
    smp.init(synth.jittered_start(d.G + 3, chains, seed=31))
    smp.sample(20, 20)                                         # warm-up of the kernels / allocator
    smp = NutsSampler(eng, seed=17)
    smp.init(synth.jittered_start(d.G + 3, chains, seed=31))
    out = smp.sample(draws, tune)
    g = np.transpose(out["trace"].cpu().numpy(), (0, 2, 1))    # [draws, C, D]
    eng.close()
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(min(n_cores, chains)) as pool:
        res = pool.map(_cfg1_cpu_chain, [(c, draws, tune, 1000) for c in range(chains)])
    cpu_wall = time.perf_counter() - t0
    c_tr = np.stack([r[0] for r in res], axis=1)               # [draws, C, D]
    cpu_leap = float(sum(r[1]["tree_size"].sum() for r in res))

    def summarise(tr):
        D = tr.shape[2]
        mean, sd = tr.mean(axis=(0, 1)), tr.std(axis=(0, 1))
        ess = np.array([st.effective_n(tr[:, :, k]) for k in range(D)])
        return mean, sd, ess

    gm, gs, ge = summarise(g)
    cm, cs, ce = summarise(c_tr)
    z = (gm - cm) / np.sqrt(gs ** 2 / ge + cs ** 2 / ce)
    gpu_secs = out["seconds_total"]
    post = out["stats"][tune:]
    return {
        "config": "synthetic 20k SNPs x 200 genes, -m normal, 4 chains, sweeps 1000 burn-in 250 (BASELINE.json config 1)",
        "gpu": {"seconds": gpu_secs, "ess_min": float(ge.min()), "ess_per_sec": float(ge.min() / gpu_secs),
                "ess_scalars": {k: float(ge[i]) for i, k in enumerate(("W_log__", "e", "mi_logodds__"))},
                "leapfrogs": float(out["stats"][:, 2].sum()), "mean_tree_accept": float(post[:, 1].mean()),
                "mean_depth": float(post[:, 0].mean()), "divergences": int(post[:, 3].sum())},
        "cpu": {"seconds": cpu_wall, "ess_min": float(ce.min()), "ess_per_sec": float(ce.min() / cpu_wall), "cores": min(n_cores, chains),
                "n_cores": n_cores, "host_cores": os.cpu_count(), "leapfrogs": cpu_leap,
                "kind": "port: oracle/nuts_oracle.py (PyMC3 3.6 NUTS semantics, numpy logp+grad), one process per chain like "
                        "pm.sample(chains=4, cores=4); not PyMC3 itself",
                "seconds_per_chain": [float(r[2]) for r in res]},
        "posterior_mean_zscore": {"max_abs": float(np.max(np.abs(z))), "rms": float(np.sqrt(np.mean(z ** 2))), "n_coordinates": int(len(z)),
                                  "note": "difference of the two posterior means of every coordinate over its combined MC standard error"},
        "ess_per_sec_ratio_gpu_over_cpu": float((ge.min() / gpu_secs) / (ce.min() / cpu_wall)),
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--with-replicas", action="store_true", help="N>1: also time chain replicas (every GPU its own 64 chains)")
    ap.add_argument("--e2e-evals", type=int, default=64, help="evaluations of the end-to-end (host buffer) measurement")
    ap.add_argument("--no-cfg1", action="store_true", help="skip the BASELINE config-1 sampler run (GPU + CPU oracle NUTS)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-sampler", action="store_true", help="skip the ESS/s section")
    ap.add_argument("--sampler-tune", type=int, default=300)
    ap.add_argument("--sampler-draws", type=int, default=200)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        if args.steps > 400:
            args.steps = 400         # a step is ~30 ms (64 chain-evals on the host threads); keep the run bounded
        bench_reference(args, wl)
    else:
        bench_gpu(args, wl)


if __name__ == "__main__":
    main()
