"""ctypes binding of libbaghera_b200.so (C ABI: include/baghera_b200.h).

The library is built in-tree by ``baghera_b200/csrc/Makefile`` (``__graft_entry__.build()``).
Loading fails loudly when it is missing: there is no eager / CPU fallback.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BGH_LIB_PATH") or os.path.join(_HERE, "libbaghera_b200.so")   # override: kernel-variant experiments (tools/)
CSRC = os.path.join(_HERE, "csrc")

BGH_ABI_VERSION = 1
BGH_OK, BGH_EINVAL, BGH_ECUDA, BGH_ENOMEM, BGH_ESTATE, BGH_ENODEV = 0, -1, -2, -3, -4, -5
BGH_MODEL_GENE_NORMAL, BGH_MODEL_GW_NORMAL, BGH_MODEL_GENE_GAMMA = 0, 1, 2


class BghError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("baghera_b200 error %d: %s" % (code, msg))
        self.code = code


class bgh_cfg(ctypes.Structure):
    _fields_ = [
        ("abi_version", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("model", ctypes.c_int32),
        ("fix_intercept", ctypes.c_int32),
        ("n_snps", ctypes.c_int64),
        ("n_genes", ctypes.c_int32),
        ("n_chains", ctypes.c_int32),
        ("c", ctypes.c_double),
        ("n_genes_total", ctypes.c_int32),
        ("first_gene_partial", ctypes.c_int32),
        ("last_gene_partial", ctypes.c_int32),
        ("first_gene_row", ctypes.c_int32),
        ("last_gene_row", ctypes.c_int32),
        ("n_shared_rows", ctypes.c_int32),
        ("sm_count_override", ctypes.c_int32),
        ("n_replicated", ctypes.c_int32),
        ("gamma_rate", ctypes.c_double),
        ("owns_replicated", ctypes.c_int32),
        ("reserved", ctypes.c_int32),
    ]


class bgh_nuts_buffers(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("state", "ckpt", "sc", "fl", "partials", "sums", "logp_w", "stats",
                                               "n_active_dev", "n_active_host")]


class bgh_nuts_cfg(ctypes.Structure):
    _fields_ = [
        ("seed", ctypes.c_uint64),
        ("target_accept", ctypes.c_double),
        ("step_scale", ctypes.c_double),
        ("emax", ctypes.c_double),
        ("da_gamma", ctypes.c_double),
        ("da_kappa", ctypes.c_double),
        ("da_t0", ctypes.c_double),
        ("n_dim_total", ctypes.c_int32),
        ("coord_offset", ctypes.c_int32),
        ("chain_offset", ctypes.c_int32),
        ("reserved", ctypes.c_int32 * 3),
    ]


class bgh_nuts_async_buffers(ctypes.Structure):
    _fields_ = [("base", bgh_nuts_buffers), ("afl", ctypes.c_void_p), ("trace", ctypes.c_void_p),
                ("run_stats", ctypes.c_void_p), ("trace_f32", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("k6", ctypes.c_void_p), ("head_trace", ctypes.c_void_p)]


NUTS_ASYNC_N_FL, NUTS_ASYNC_N_SC, NUTS_ASYNC_PARTIALS = 26, 24, 24
NUTS_MAX_DEPTH, NUTS_N_SLOTS, NUTS_N_SC, NUTS_N_FL, NUTS_N_STATS = 10, 21, 16, 11, 8
NUTS_MAX_PARTIALS = 1 + 2 * NUTS_MAX_DEPTH
NUTS_SLOT_Q, NUTS_SLOT_GRAD, NUTS_SLOT_VAR = 0, 1, 2
NUTS_SLOT_FG_MEAN, NUTS_SLOT_FG_RAW, NUTS_SLOT_BG_MEAN, NUTS_SLOT_BG_RAW = 16, 17, 18, 19
NUTS_SC_EPS, NUTS_SC_LOGP = 0, 1
NUTS_AF_BAD_START = 21      # kAfBadStart (csrc/bgh_sampler.h): the energy at the start of a transition was not finite
NUTS_AF_TUNE_END_US = NUTS_ASYNC_N_FL - 1     # kAfTuneEndUs: %globaltimer >> 10 (low 32 bits) at the end of the chain's last tuning transition
NUTS_SC_RUN_START_US = NUTS_ASYNC_N_SC - 1    # kScRunStartUs: %globaltimer >> 10 at the start of the run

_vp = ctypes.c_void_p
_i32, _i64, _dbl = ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_pi32 = ctypes.POINTER(ctypes.c_int32)
_pi64 = ctypes.POINTER(ctypes.c_int64)
_pdbl = ctypes.POINTER(ctypes.c_double)

# name -> (restype, argtypes); mirrors include/baghera_b200.h one to one
SIGNATURES = {
    "bgh_cfg_default": (None, [ctypes.POINTER(bgh_cfg)]),
    "bgh_version": (ctypes.c_char_p, []),
    "bgh_create": (_i32, [ctypes.POINTER(_vp), ctypes.POINTER(bgh_cfg)]),
    "bgh_destroy": (None, [_vp]),
    "bgh_last_error": (ctypes.c_char_p, [_vp]),
    "bgh_dim": (_i32, [_vp]),
    "bgh_chain_stride": (_i32, [_vp]),
    "bgh_payload_rows": (_i32, [_vp]),
    "bgh_upload": (_i32, [_vp, _vp, _vp, _vp]),
    "bgh_lik_const": (_dbl, [_vp]),
    "bgh_set_lik_const": (_i32, [_vp, _dbl]),
    "bgh_logp_grad": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "bgh_logp_grad_local": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "bgh_logp_grad_finish": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "bgh_peer_handle_bytes": (_i32, []),
    "bgh_peer_export": (_i32, [_vp, _vp]),
    "bgh_peer_connect": (_i32, [_vp, _i32, _i32, _vp]),
    "bgh_logp_grad_sharded": (_i32, [_vp, _vp, _vp, _vp, _vp]),
    "bgh_peer_status": (_i32, [_vp, _pi32]),
    "bgh_peer_connect_local": (_i32, [_vp, _i32, _i32, ctypes.POINTER(_vp)]),
    "bgh_debug_peer_mode": (_i32, [_vp, _i32, _i32]),
    "bgh_logp_grad_host": (_i32, [_vp, _vp, _vp, _vp]),
    "bgh_logp_grad_host_stream": (_i32, [_vp, _i32, _vp, _vp, _vp]),
    "bgh_to_device_layout": (_i32, [_vp, _vp, _vp, _vp]),
    "bgh_from_device_layout": (_i32, [_vp, _vp, _vp, _vp]),
    "bgh_n_groups": (_i32, [_vp]),
    "bgh_get_partition": (_i32, [_vp, _pi32, _i32]),
    "bgh_get_sampler_partition": (_i32, [_vp, _pi32, _i32]),
    "bgh_n_split_genes": (_i32, [_vp]),
    "bgh_stream_len": (_i64, [_vp]),
    "bgh_debug_stream": (_i32, [_vp, _i32, _vp, _i64]),
    "bgh_launch_count": (_i64, [_vp]),
    "bgh_set_profiling": (_i32, [_vp, _i32]),
    "bgh_kernel_time_ms": (_i32, [_vp, _pdbl, _pdbl, _pi64, _i32]),
    "bgh_dfma_peak": (_i32, [_vp, _pdbl]),
    "bgh_nuts_cfg_default": (None, [ctypes.POINTER(bgh_nuts_cfg)]),
    "bgh_nuts_vec_blocks": (_i32, [_vp]),
    "bgh_nuts_init": (_i32, [_vp, ctypes.POINTER(bgh_nuts_buffers), ctypes.POINTER(bgh_nuts_cfg), _vp]),
    "bgh_nuts_transition": (_i32, [_vp, ctypes.POINTER(bgh_nuts_buffers), ctypes.POINTER(bgh_nuts_cfg), _i64, _i32, _i32,
                                   _dbl, _dbl, _i32, _vp, _pi32]),
    "bgh_set_row_coords": (_i32, [_vp, _pi32]),
    "bgh_nuts_stop_tuning": (_i32, [_vp, ctypes.POINTER(bgh_nuts_buffers), ctypes.POINTER(bgh_nuts_cfg), _vp]),
    "bgh_nuts_run": (_i32, [_vp, ctypes.POINTER(bgh_nuts_async_buffers), ctypes.POINTER(bgh_nuts_cfg), _i32, _i32, _i32,
                            _vp, _pi64]),
    "bgh_debug_trace": (_i32, [_vp, _i32, _vp, _i64]),
    "bgh_debug_sampler_plan": (_i32, [_i64, _pi32, _i32, _i32, _i32, _i32, _pi32, _pi32]),
    "bgh_csv_read": (_i32, [ctypes.c_char_p, _i32, _i32, ctypes.POINTER(_vp)]),
    "bgh_ingest_last_error": (ctypes.c_char_p, []),
    "bgh_table_rows": (_i64, [_vp]),
    "bgh_table_n_columns": (_i32, [_vp]),
    "bgh_table_column_name": (ctypes.c_char_p, [_vp, _i32]),
    "bgh_table_column_kind": (_i32, [_vp, _i32]),
    "bgh_table_f64": (_vp, [_vp, _i32]),
    "bgh_table_codes": (_vp, [_vp, _i32]),
    "bgh_table_n_labels": (_i32, [_vp, _i32]),
    "bgh_table_label": (ctypes.c_char_p, [_vp, _i32, _i32]),
    "bgh_table_free": (None, [_vp]),
    "bgh_debug_plan": (_i32, [_i64, _pi32, _i32, _i32, _i32, _i32, _i32, _pi32, _pi32, _pi32, _pi32, _pi32, _pi32, _i32]),
}

_LIB = None


def build(verbose: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    out = subprocess.run(["make", "-j4", "-C", CSRC], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libbaghera_b200.so failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)
    return LIB_PATH


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libbaghera_b200.so is missing (%s). Build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` or `make -C baghera_b200/csrc`. There is no CPU fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def check(rc: int, ctx=None):
    if rc != BGH_OK:
        msg = lib().bgh_last_error(ctx)
        raise BghError(rc, msg.decode() if msg else "?")
