"""PyTorch-hosted batched-chain NUTS driver (replaces ``pm.sample`` at
ref: baghera_tool/gene_regression.py:100-106 and ref: baghera_tool/heritability.py:56-62).

torch owns the memory and the stream; every arithmetic step of a transition (momentum refresh,
leapfrog, likelihood + gradient, tree bookkeeping, accept / multinomial draws, adaptation) runs in
libbaghera_b200.so (csrc/bgh_sampler.cu, csrc/bgh_lik.cu).  Semantics: PyMC3 3.6 NUTS with
``init='jitter+adapt_diag'`` [recalled, SURVEY §8a] — one independent adaptation per chain, tuning
draws discarded, ``early_max_treedepth=8`` for the first 200 tuning draws then ``max_treedepth=10``.
"""
from __future__ import annotations

import ctypes
import time

import numpy as np

from . import _lib
from ._lib import (NUTS_ASYNC_N_FL, NUTS_ASYNC_N_SC, NUTS_ASYNC_PARTIALS, bgh_nuts_async_buffers, NUTS_MAX_DEPTH, NUTS_MAX_PARTIALS, NUTS_N_FL, NUTS_N_SC, NUTS_N_SLOTS, NUTS_N_STATS,
                   NUTS_SC_EPS, NUTS_SC_LOGP, NUTS_SLOT_BG_MEAN, NUTS_SLOT_BG_RAW, NUTS_SLOT_FG_MEAN,
                   NUTS_SLOT_FG_RAW, NUTS_SLOT_GRAD, NUTS_SLOT_Q, NUTS_SLOT_VAR, bgh_nuts_buffers, bgh_nuts_cfg,
                   check)

STAT_NAMES = ("depth", "mean_tree_accept", "tree_size", "diverging", "energy", "step_size", "model_logp",
              "max_energy_error")
ADAPT_WINDOW = 101          # quadpotential.QuadPotentialDiagAdapt(adaptation_window=101)
INITIAL_WEIGHT = 10.0       # init_nuts('jitter+adapt_diag'): QuadPotentialDiagAdapt(ndim, mean, var, 10)
EARLY_MAX_TREEDEPTH, MAX_TREEDEPTH, EARLY_DRAWS = 8, 10, 200


class NutsSampler:
    def __init__(self, engine, seed=20211, target_accept=0.90, row_coords=None, n_dim_total=0, chain_offset=0):
        """row_coords / n_dim_total: gene-block sharded engines only (``Shard.row_coords()`` and the global D): every
        rank runs its own NutsSampler over its rows of q; the per-chain sums are exchanged over NVLink peer memory
        inside the sampler's kernels and the momentum normals are keyed by the global coordinate, so all ranks take
        identical decisions and reproduce the single-GPU chains."""
        import torch

        self.eng = engine
        self.t = torch
        self.L = engine.L
        dev, f64 = engine.device, torch.float64
        D, Cp = engine.D, engine.Cp
        self.n_vec_blocks = int(self.L.bgh_nuts_vec_blocks(engine.ctx))
        self.state = torch.zeros((NUTS_N_SLOTS, D, Cp), dtype=f64, device=dev)
        self.ckpt = torch.zeros((2, NUTS_MAX_DEPTH, D, Cp), dtype=f64, device=dev)
        self.sc = torch.zeros((max(NUTS_N_SC, NUTS_ASYNC_N_SC), Cp), dtype=f64, device=dev)
        self.afl = torch.zeros((NUTS_ASYNC_N_FL, Cp), dtype=torch.int32, device=dev)
        self.fl = torch.zeros((NUTS_N_FL, Cp), dtype=torch.int32, device=dev)
        nq = max(NUTS_MAX_PARTIALS, NUTS_ASYNC_PARTIALS)
        self.partials = torch.zeros((self.n_vec_blocks, nq, Cp), dtype=f64, device=dev)
        self.sums = torch.zeros((nq, Cp), dtype=f64, device=dev)
        self.logp_w = torch.zeros(Cp, dtype=f64, device=dev)
        self.stats = torch.zeros((NUTS_N_STATS, Cp), dtype=f64, device=dev)
        self.n_active_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self.n_active_host = torch.zeros(1, dtype=torch.int32).pin_memory()
        b = bgh_nuts_buffers()
        for name in ("state", "ckpt", "sc", "fl", "partials", "sums", "logp_w", "stats"):
            setattr(b, name, getattr(self, name).data_ptr())
        b.n_active_dev = self.n_active_dev.data_ptr()
        b.n_active_host = self.n_active_host.data_ptr()
        self.buf = b
        cfg = bgh_nuts_cfg()
        self.L.bgh_nuts_cfg_default(ctypes.byref(cfg))
        cfg.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        cfg.target_accept = float(target_accept)
        cfg.coord_offset = 0
        cfg.n_dim_total = int(n_dim_total)
        cfg.chain_offset = int(chain_offset)      # chains dealt to several devices: Philox keys by the global chain index
        if row_coords is not None:
            engine.set_row_coords(row_coords)
        self.cfg = cfg
        self.draw = 0
        self.n_adapt = 0
        self.fg_w, self.bg_w = INITIAL_WEIGHT, 0.0
        self.total_leapfrogs = 0

    # ------------------------------------------------------------------ views
    @property
    def q(self):
        return self.state[NUTS_SLOT_Q]

    @property
    def var(self):
        return self.state[NUTS_SLOT_VAR]

    @property
    def step_size(self):
        return self.sc[NUTS_SC_EPS, :self.eng.C]

    @property
    def logp(self):
        return self.sc[NUTS_SC_LOGP, :self.eng.C]

    def _stream(self):
        return ctypes.c_void_p(self.t.cuda.current_stream(self.eng.device).cuda_stream)

    # ------------------------------------------------------------------ init='jitter+adapt_diag'
    def init(self, q0_cd, start_mean=None):
        """q0_cd: float64[C, D] start points (test point + U(-1,1) jitter per chain).  start_mean: mean of the start points
        of ALL chains of the analysis when this sampler runs a block of them (jitter+adapt_diag seeds the mass-matrix
        estimator with it); default: the mean of q0_cd."""
        t = self.t
        q0 = np.asarray(q0_cd, dtype=np.float64)
        self.state[NUTS_SLOT_Q].copy_(self.eng.to_device_layout(q0))
        self.state[NUTS_SLOT_VAR].fill_(1.0)
        mean = t.as_tensor(q0.mean(axis=0) if start_mean is None else np.asarray(start_mean, dtype=np.float64),
                           dtype=t.float64, device=self.eng.device)
        self.state[NUTS_SLOT_FG_MEAN].copy_(mean[:, None].expand(-1, self.eng.Cp))
        self.state[NUTS_SLOT_FG_RAW].fill_(INITIAL_WEIGHT)          # raw_var = initial_diag * weight
        self.state[NUTS_SLOT_BG_MEAN].zero_()
        self.state[NUTS_SLOT_BG_RAW].zero_()
        self.fg_w, self.bg_w, self.n_adapt, self.draw = INITIAL_WEIGHT, 0.0, 0, 0
        check(self.L.bgh_nuts_init(self.eng.ctx, ctypes.byref(self.buf), ctypes.byref(self.cfg), self._stream()),
              self.eng.ctx)
        t.cuda.synchronize(self.eng.device)
        lp = self.logp.cpu().numpy()
        if not np.all(np.isfinite(lp)):
            raise ValueError("Bad initial energy: logp is not finite at the start point of chains %s"
                             % np.flatnonzero(~np.isfinite(lp)).tolist())

    # ------------------------------------------------------------------ one transition
    def transition(self, tune, max_treedepth=MAX_TREEDEPTH):
        fg_w = bg_w = 0.0
        switch = 0
        if tune:
            # quadpotential.QuadPotentialDiagAdapt.update: add the sample to both estimators, refresh
            # the variance from the foreground one, switch windows every ADAPT_WINDOW samples
            self.fg_w += 1.0
            self.bg_w += 1.0
            fg_w, bg_w = self.fg_w, self.bg_w
            switch = int(self.n_adapt > 0 and self.n_adapt % ADAPT_WINDOW == 0)
        n_leap = ctypes.c_int32(0)
        check(self.L.bgh_nuts_transition(self.eng.ctx, ctypes.byref(self.buf), ctypes.byref(self.cfg), self.draw,
                                         int(bool(tune)), int(max_treedepth), fg_w, bg_w, switch, self._stream(),
                                         ctypes.byref(n_leap)), self.eng.ctx)
        if tune:
            if switch:
                self.fg_w, self.bg_w = self.bg_w, 0.0
            self.n_adapt += 1
        self.draw += 1
        self.total_leapfrogs += n_leap.value
        return n_leap.value

    def stop_tuning(self):
        check(self.L.bgh_nuts_stop_tuning(self.eng.ctx, ctypes.byref(self.buf), ctypes.byref(self.cfg), self._stream()),
              self.eng.ctx)

    # ------------------------------------------------------------------ pm.sample, asynchronous chains
    def sample(self, draws, tune, trace_dtype=None, progress=None, poll_ticks=64, stream_stats=False, keep_trace=True):
        """pm.sample(draws, tune=tune): the whole run on the device (bgh_nuts_run).  Every tick advances
        every chain by one leapfrog; chains never wait for the deepest tree of the batch.  Returns the
        same dict as :meth:`sample_lockstep` (``leapfrogs`` holds the per-draw mean tree size over chains).
        ``progress`` is accepted for signature compatibility with :meth:`sample_lockstep` and ignored: the run is ONE blocking
        C call whose host loop only polls the number of finished chains every ``poll_ticks`` ticks.

        stream_stats: the kernel also accumulates, per coordinate and chain, {sum, sum of squares, count(value - mi > 0)} of
        the recorded per-gene effects in fp64 (``k6`` [3, D, C]) and keeps the shared parameters of every draw in fp64
        (``head_trace`` [draws, H, C]) — the per-gene table of ref: gene_regression.py:165-174 then needs the stored positions
        only for its quantiles (an fp32 trace is enough) or not at all (keep_trace=False)."""
        t = self.t
        eng = self.eng
        C = eng.C
        dt = trace_dtype or t.float64
        if dt not in (t.float64, t.float32):
            raise ValueError("trace dtype must be float64 or float32")
        # the device writes the trace chain-major [draws, C, D] (contiguous per chain); callers see the
        # reference-shaped view [draws, D, C]
        trace_cd = t.empty((max(draws, 1), C, eng.D), dtype=dt, device=eng.device) if keep_trace else None
        trace = trace_cd.permute(0, 2, 1) if keep_trace else None
        H = 3 if eng.model == "gene" else 2
        k6 = t.zeros((3, eng.D, eng.Cp), dtype=t.float64, device=eng.device) if stream_stats else None
        head = t.zeros((max(draws, 1), C, H), dtype=t.float64, device=eng.device) if stream_stats else None
        stats = t.zeros((tune + draws, NUTS_N_STATS, C), dtype=t.float64, device=eng.device)
        ab = bgh_nuts_async_buffers()
        ab.base = self.buf
        ab.afl = self.afl.data_ptr()
        ab.trace = trace_cd.data_ptr() if keep_trace else None
        ab.k6 = k6.data_ptr() if stream_stats else None
        ab.head_trace = head.data_ptr() if stream_stats else None
        ab.run_stats = stats.data_ptr()
        ab.trace_f32 = int(dt == t.float32)
        n_ticks = ctypes.c_int64(0)
        t.cuda.synchronize(eng.device)
        t0 = time.perf_counter()
        check(self.L.bgh_nuts_run(eng.ctx, ctypes.byref(ab), ctypes.byref(self.cfg), int(tune), int(draws),
                                  int(poll_ticks), self._stream(), ctypes.byref(n_ticks)), eng.ctx)
        t.cuda.synchronize(eng.device)
        total = time.perf_counter() - t0
        if int(self.afl[_lib.NUTS_AF_BAD_START, :C].sum().item()) != 0:
            raise ValueError("Bad initial energy: the energy at the start of a transition was not finite")
        st = stats.cpu().numpy()
        tree = st[:, 2, :]                                   # leapfrogs per transition and chain
        frac_tune = tree[:tune].sum() / max(tree.sum(), 1.0)
        seconds_tune, note = total * frac_tune, "seconds_total is measured; tune / draws are apportioned by the leapfrogs of each phase"
        if tune > 0:
            # measured: device timestamps (units of 1.024 us) at the start of the run and at the end of every chain's last tuning
            # transition (chains leave tuning at different ticks: the mean over chains is the time the batch spent tuning)
            t_end = self.afl[_lib.NUTS_AF_TUNE_END_US, :C].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            t_start = self.sc[_lib.NUTS_SC_RUN_START_US, :C].cpu().numpy().astype(np.int64) & 0xFFFFFFFF
            dt = ((t_end - t_start) & 0xFFFFFFFF) * 1.024e-6
            if np.all(dt > 0.0) and np.all(dt <= total * 1.05 + 1e-3):
                seconds_tune = float(min(dt.mean(), total)) if draws > 0 else total
                note = ("seconds_total is wall clock; seconds_tune is measured on the device (globaltimer at the end of each chain's "
                        "last tuning transition, mean over chains), seconds_draws the rest")
        self.total_leapfrogs += int(n_ticks.value)
        self.draw += tune + draws
        out = dict(trace=trace[:draws] if keep_trace else None, stats=st, leapfrogs=tree.mean(axis=1), ticks=int(n_ticks.value),
                   seconds_tune=seconds_tune, seconds_draws=total - seconds_tune, seconds_total=total, tune=tune,
                   seconds_note=note)
        if stream_stats:
            out["k6"] = k6[:, :, :C]
            out["head_trace"] = head[:draws].permute(0, 2, 1)          # [draws, H, C]
        return out

    # ------------------------------------------------------------------ pm.sample, lock-step chains
    def sample_lockstep(self, draws, tune, trace_dtype=None, progress=None):
        """Runs tune + draws transitions; returns a dict with
        ``trace``  torch tensor [draws, D, C] on the device (positions in the unconstrained space),
        ``stats``  numpy float64 [tune + draws, 8, C] (STAT_NAMES), ``leapfrogs`` int64 [tune + draws],
        ``seconds_tune`` / ``seconds_draws`` wall-clock, ``tune``."""
        t = self.t
        eng = self.eng
        C = eng.C
        dt = trace_dtype or t.float64
        trace = t.empty((draws, eng.D, C), dtype=dt, device=eng.device)
        stats = t.empty((tune + draws, NUTS_N_STATS, C), dtype=t.float64, device=eng.device)
        leap = np.zeros(tune + draws, dtype=np.int64)
        t.cuda.synchronize(eng.device)
        t0 = time.perf_counter()
        t_tune = 0.0
        for i in range(tune + draws):
            tuning = i < tune
            if i == tune:
                self.stop_tuning()
                t.cuda.synchronize(eng.device)
                t_tune = time.perf_counter() - t0
            depth = EARLY_MAX_TREEDEPTH if (tuning and i < EARLY_DRAWS) else MAX_TREEDEPTH
            leap[i] = self.transition(tuning, depth)
            stats[i].copy_(self.stats[:, :C])
            if not tuning:
                trace[i - tune].copy_(self.q[:, :C])
            if progress is not None and (i + 1) % progress == 0:
                print("  draw %d/%d  leapfrogs/draw %.1f" % (i + 1, tune + draws, leap[max(0, i - progress + 1):i + 1].mean()),
                      flush=True)
        t.cuda.synchronize(eng.device)
        total = time.perf_counter() - t0
        if tune >= tune + draws:
            t_tune = total
        st = stats.cpu().numpy()
        if np.any(st[:, 3] > 0):
            pass   # divergences are reported by the caller (PyMC3 prints a warning)
        return dict(trace=trace, stats=st, leapfrogs=leap, seconds_tune=t_tune, seconds_draws=total - t_tune, tune=tune)


def smoke(engine, data):
    """A few tuned transitions on a tiny problem (called by __graft_entry__.smoke)."""
    from . import synth
    s = NutsSampler(engine, seed=1)
    s.init(synth.jittered_start(engine.D, engine.C, seed=2))
    for i in range(6):
        s.transition(True, 6)
    engine.torch.cuda.synchronize()
    lp = s.logp.cpu().numpy()
    assert np.all(np.isfinite(lp)), lp
    return lp
