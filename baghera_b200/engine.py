"""Device engine: owns one ``bgh_ctx`` (one GPU / rank) and exposes logp+grad on torch tensors.

PyTorch is plumbing here (device memory, streams, torch.distributed); all arithmetic of the path
runs in libbaghera_b200.so.  Replaces ``model.logp_dlogp_function()`` of the PyMC3 model built at
ref: baghera_tool/gene_regression.py:77-98 / baghera_tool/heritability.py:44-55.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import BGH_MODEL_GENE_GAMMA, BGH_MODEL_GENE_NORMAL, BGH_MODEL_GW_NORMAL, BghError, bgh_cfg, check


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class Engine:
    """One context per GPU.  Parameters live on the device as float64[D, Cp] (gene-major, chain-minor)."""

    def __init__(self, chi2, ld, gene_id, n_genes, c, n_chains, model="gene", fix_intercept=False,
                 device=0, shard=None, gamma_rate=None):
        """model: "gene" (normal prior, ref gene_regression.py:77-98), "gamma" (ref :229-248; pass
        c = n_patients and gamma_rate = N_1kG) or "gw" (ref heritability.py:44-55)."""
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError("baghera_b200.Engine needs a CUDA device (sm_100a); there is no CPU fallback")
        self.torch = torch
        self.L = _lib.lib()
        chi2 = np.ascontiguousarray(chi2, dtype=np.float32)
        ld = np.ascontiguousarray(ld, dtype=np.float32)
        cfg = bgh_cfg()
        self.L.bgh_cfg_default(ctypes.byref(cfg))
        cfg.device = int(device)
        if model not in ("gene", "gamma", "gw"):
            raise ValueError("unknown model %r" % (model,))
        cfg.model = {"gene": BGH_MODEL_GENE_NORMAL, "gamma": BGH_MODEL_GENE_GAMMA, "gw": BGH_MODEL_GW_NORMAL}[model]
        if model == "gamma":
            if gamma_rate is None:
                raise ValueError("the gamma model needs gamma_rate (N_1kG)")
            cfg.gamma_rate = float(gamma_rate)
        cfg.fix_intercept = int(bool(fix_intercept))
        cfg.n_snps = int(chi2.shape[0])
        cfg.n_genes = int(n_genes) if model != "gw" else 1
        cfg.n_chains = int(n_chains)
        cfg.c = float(c)
        if shard is not None:
            for k, v in shard.items():
                setattr(cfg, k, int(v))
        self.cfg = cfg
        self.model = model
        self.device = torch.device("cuda", int(device))
        ctx = ctypes.c_void_p()
        check(self.L.bgh_create(ctypes.byref(ctx), ctypes.byref(cfg)))
        self.ctx = ctx
        gid = None
        if model != "gw":
            gid = np.ascontiguousarray(gene_id, dtype=np.int32)
            if gid.shape[0] != chi2.shape[0]:
                raise ValueError("gene_id length mismatch")
        check(self.L.bgh_upload(ctx, chi2.ctypes.data_as(ctypes.c_void_p), ld.ctypes.data_as(ctypes.c_void_p),
                                gid.ctypes.data_as(ctypes.c_void_p) if gid is not None else None), ctx)
        self.M = int(chi2.shape[0])
        self.C = int(n_chains)
        self.D = int(self.L.bgh_dim(ctx))
        self.Cp = int(self.L.bgh_chain_stride(ctx))
        self.payload_rows = int(self.L.bgh_payload_rows(ctx))
        self.lik_const = float(self.L.bgh_lik_const(ctx))

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "ctx", None):
            self.L.bgh_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return ctypes.c_void_p(self.torch.cuda.current_stream(self.device).cuda_stream)

    def empty_params(self):
        return self.torch.empty((self.D, self.Cp), dtype=self.torch.float64, device=self.device)

    def to_device_layout(self, q_cd):
        """float64[C, D] (host numpy or device tensor, PyMC3 chain-major) -> device float64[D, Cp]."""
        t = self.torch
        q = t.as_tensor(np.asarray(q_cd, dtype=np.float64) if not t.is_tensor(q_cd) else q_cd).to(
            self.device, t.float64).contiguous()
        assert q.shape == (self.C, self.D), (q.shape, (self.C, self.D))
        out = self.empty_params()
        check(self.L.bgh_to_device_layout(self.ctx, _ptr(q), _ptr(out), self._stream()), self.ctx)
        return out

    def from_device_layout(self, x_dc):
        t = self.torch
        out = t.empty((self.C, self.D), dtype=t.float64, device=self.device)
        check(self.L.bgh_from_device_layout(self.ctx, _ptr(x_dc), _ptr(out), self._stream()), self.ctx)
        return out

    # ------------------------------------------------------------------ the hot path
    def logp_grad(self, q_dc, logp=None, grad=None):
        """Async on the current torch stream. q_dc: float64[D, Cp] -> (logp[Cp], grad[D, Cp])."""
        t = self.torch
        if logp is None:
            logp = t.empty(self.Cp, dtype=t.float64, device=self.device)
        if grad is None:
            grad = t.empty_like(q_dc)
        check(self.L.bgh_logp_grad(self.ctx, _ptr(q_dc), _ptr(logp), _ptr(grad), self._stream()), self.ctx)
        return logp, grad

    def logp_grad_local(self, q_dc, grad, payload):
        check(self.L.bgh_logp_grad_local(self.ctx, _ptr(q_dc), _ptr(grad), _ptr(payload), self._stream()), self.ctx)

    def logp_grad_finish(self, q_dc, payload, logp, grad):
        check(self.L.bgh_logp_grad_finish(self.ctx, _ptr(q_dc), _ptr(payload), _ptr(logp), _ptr(grad),
                                          self._stream()), self.ctx)

    # ------------------------------------------------------------------ fused peer-memory exchange
    def peer_export(self):
        """Allocates this rank's mailbox; returns its CUDA IPC handle (bytes)."""
        n = int(self.L.bgh_peer_handle_bytes())
        buf = (ctypes.c_ubyte * n)()
        check(self.L.bgh_peer_export(self.ctx, ctypes.cast(buf, ctypes.c_void_p)), self.ctx)
        return bytes(buf)

    def peer_connect(self, rank, world, handles):
        """handles: list of `world` IPC handles (bytes) in rank order (all-gathered by the caller)."""
        blob = b"".join(handles)
        buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        check(self.L.bgh_peer_connect(self.ctx, int(rank), int(world), ctypes.cast(buf, ctypes.c_void_p)), self.ctx)

    def peer_connect_local(self, rank, engines):
        """Ranks as engines of this process (tests emulating the ranks on one GPU; one process driving several GPUs)."""
        arr = (ctypes.c_void_p * len(engines))(*[e.ctx for e in engines])
        check(self.L.bgh_peer_connect_local(self.ctx, int(rank), len(engines), arr), self.ctx)

    def debug_peer_mode(self, send_only, seq=-1):
        check(self.L.bgh_debug_peer_mode(self.ctx, int(bool(send_only)), int(seq)), self.ctx)

    def set_row_coords(self, coords):
        c = np.ascontiguousarray(coords, dtype=np.int32)
        assert c.shape == (self.D,), (c.shape, self.D)
        check(self.L.bgh_set_row_coords(self.ctx, c.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))), self.ctx)

    def logp_grad_sharded(self, q_dc, logp, grad):
        """One launch per rank: local pass + one-shot all-reduce over peer memory + closing formulas."""
        check(self.L.bgh_logp_grad_sharded(self.ctx, _ptr(q_dc), _ptr(logp), _ptr(grad), self._stream()), self.ctx)

    def peer_status(self):
        v = ctypes.c_int32()
        check(self.L.bgh_peer_status(self.ctx, ctypes.byref(v)), self.ctx)
        return v.value

    def logp_grad_host(self, q_cd, logp=None, grad=None, n_evals=None):
        """B1' with host buffers: q float64[C, D] (or [n, C, D]) numpy -> (logp, grad) numpy."""
        q = np.ascontiguousarray(q_cd, dtype=np.float64)
        batched = q.ndim == 3
        n = q.shape[0] if batched else 1
        assert q.shape[-2:] == (self.C, self.D)
        if logp is None:
            logp = np.empty((n, self.C) if batched else (self.C,))
        if grad is None:
            grad = np.empty_like(q)
        check(self.L.bgh_logp_grad_host_stream(self.ctx, n, q.ctypes.data_as(ctypes.c_void_p),
                                               logp.ctypes.data_as(ctypes.c_void_p),
                                               grad.ctypes.data_as(ctypes.c_void_p)), self.ctx)
        return logp, grad

    def logp_grad_host_ptr(self, n, q_ptr, logp_ptr, grad_ptr):
        check(self.L.bgh_logp_grad_host_stream(self.ctx, int(n), ctypes.c_void_p(q_ptr), ctypes.c_void_p(logp_ptr),
                                               ctypes.c_void_p(grad_ptr)), self.ctx)

    def set_lik_const(self, total):
        """Sharded runs: install the all-reduced normalising constant (sum over ranks of lik_const)."""
        check(self.L.bgh_set_lik_const(self.ctx, float(total)), self.ctx)
        self.lik_const = float(total)

    # ------------------------------------------------------------------ introspection
    def partition(self, sampler=False):
        """SNP-range offsets (n_groups + 1) of the evaluation plan, or of the sampler kernel's plan."""
        n = int(self.L.bgh_n_groups(self.ctx)) + 1
        off = np.empty(n, dtype=np.int32)
        fn = self.L.bgh_get_sampler_partition if sampler else self.L.bgh_get_partition
        check(fn(self.ctx, off.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), n), self.ctx)
        return off

    def device_stream(self):
        """The gene-sorted, padded stream bgh_upload built on the device: (s/sqrt(l), 1/sqrt(l), sqrt(l), gene id)."""
        n = int(self.L.bgh_stream_len(self.ctx))
        out = []
        for which in range(4):
            a = np.empty(n, dtype=np.int32 if which == 3 else np.float64)
            check(self.L.bgh_debug_stream(self.ctx, which, a.ctypes.data_as(ctypes.c_void_p), n), self.ctx)
            out.append(a)
        return tuple(out)

    @property
    def n_split_genes(self):
        return int(self.L.bgh_n_split_genes(self.ctx))

    @property
    def launch_count(self):
        return int(self.L.bgh_launch_count(self.ctx))

    def set_profiling(self, on=True):
        check(self.L.bgh_set_profiling(self.ctx, int(bool(on))), self.ctx)

    def kernel_time_ms(self, reset=True):
        a, b, n = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
        check(self.L.bgh_kernel_time_ms(self.ctx, ctypes.byref(a), ctypes.byref(b), ctypes.byref(n), int(reset)), self.ctx)
        return a.value, b.value, n.value

    def dfma_peak(self):
        v = ctypes.c_double()
        check(self.L.bgh_dfma_peak(self.ctx, ctypes.byref(v)), self.ctx)
        return v.value


def debug_plan(gid_sorted, G, n_cta, groups_per_cta=16, first_partial=False, last_partial=False):
    """Host-only work plan (no GPU): returns dict(off, flags, split_gene, split_ptr, split_slots, n_cta)."""
    n_groups = n_cta * groups_per_cta
    L = _lib.lib()
    gid = np.ascontiguousarray(gid_sorted, dtype=np.int32)
    M = gid.shape[0]
    cap = 3 * n_groups + 8
    off = np.zeros(n_groups + 1, np.int32)
    flags = np.zeros(n_groups, np.int32)
    ns = ctypes.c_int32()
    sg = np.zeros(cap, np.int32)
    sp = np.zeros(cap + 1, np.int32)
    ss = np.zeros(cap, np.int32)
    P = ctypes.POINTER(ctypes.c_int32)
    rc = L.bgh_debug_plan(M, gid.ctypes.data_as(P), int(G), int(n_cta), int(groups_per_cta), int(first_partial), int(last_partial),
                          off.ctypes.data_as(P), flags.ctypes.data_as(P), ctypes.byref(ns), sg.ctypes.data_as(P),
                          sp.ctypes.data_as(P), ss.ctypes.data_as(P), cap)
    check(rc)
    n = ns.value
    return dict(off=off, flags=flags, n_cta=n_cta, groups_per_cta=groups_per_cta, split_gene=sg[:n].copy(), split_ptr=sp[:n + 1].copy(),
                split_slots=ss[:sp[n]].copy())


def debug_sampler_plan(gid_sorted, G, n_cta, groups_per_cta=16, two_level=True):
    """Host-only sampler work plan (bgh_upload's defaults): returns (offsets[n_groups + 1], flags[n_groups])."""
    n_groups = n_cta * groups_per_cta
    L = _lib.lib()
    gid = np.ascontiguousarray(gid_sorted, dtype=np.int32)
    off = np.zeros(n_groups + 1, np.int32)
    flags = np.zeros(n_groups, np.int32)
    P = ctypes.POINTER(ctypes.c_int32)
    check(L.bgh_debug_sampler_plan(gid.shape[0], gid.ctypes.data_as(P), int(G), int(n_cta), int(groups_per_cta), int(bool(two_level)),
                                   off.ctypes.data_as(P), flags.ctypes.data_as(P)))
    return off, flags
