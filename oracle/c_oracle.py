"""ctypes loader for oracle/liboracle.so (the C restatement in oracle/oracle_c.c).

TEST INFRASTRUCTURE / REPORTED CPU BASELINE ONLY — see the header of oracle_c.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int64)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle_c.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        L = ctypes.CDLL(so)
        L.bgh_oracle_lik_const.restype = ctypes.c_double
        L.bgh_oracle_lik_const.argtypes = [ctypes.c_int64, _dp]
        L.bgh_oracle_logp_grad.restype = ctypes.c_int
        L.bgh_oracle_logp_grad.argtypes = [ctypes.c_int64, ctypes.c_int64, _dp, _dp, _ip, ctypes.c_double,
                                           ctypes.c_double, ctypes.c_int, _dp, _dp, _dp, _dp]
        L.bgh_oracle_gw_logp_grad.restype = ctypes.c_int
        L.bgh_oracle_gw_logp_grad.argtypes = [ctypes.c_int64, _dp, _dp, ctypes.c_double, ctypes.c_double,
                                              _dp, _dp, _dp, _dp]
        L.bgh_oracle_logp_grad_batch.restype = ctypes.c_int
        L.bgh_oracle_logp_grad_batch.argtypes = [ctypes.c_int64, ctypes.c_int64, _dp, _dp, _ip,
                                                 ctypes.c_double, ctypes.c_double, ctypes.c_int,
                                                 ctypes.c_int64, _dp, _dp, _dp, ctypes.c_int]
        L.bgh_oracle_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _p(a, t=_dp):
    return a.ctypes.data_as(t)


class COracle:
    """Holds fp64 copies of the inputs in FILE order (unsorted cat), like the reference."""

    def __init__(self, chi2, ld, cat, G, c, fix_intercept=False):
        self.chi2 = np.ascontiguousarray(chi2, dtype=np.float64)
        self.ld = np.ascontiguousarray(ld, dtype=np.float64)
        self.cat = np.ascontiguousarray(cat, dtype=np.int64)
        self.M, self.G, self.c = int(self.chi2.shape[0]), int(G), float(c)
        self.fix = int(bool(fix_intercept))
        self.const = lib().bgh_oracle_lik_const(self.M, _p(self.ld))
        self._work = np.empty(2 * self.M)

    def logp_grad(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        lp = np.empty(1)
        g = np.empty(self.G + 3)
        lib().bgh_oracle_logp_grad(self.M, self.G, _p(self.chi2), _p(self.ld), _p(self.cat, _ip), self.c,
                                   self.const, self.fix, _p(q), _p(lp), _p(g), _p(self._work))
        return float(lp[0]), g

    def gw_logp_grad(self, q):
        q = np.ascontiguousarray(q, dtype=np.float64)
        lp = np.empty(1)
        g = np.empty(2)
        lib().bgh_oracle_gw_logp_grad(self.M, _p(self.chi2), _p(self.ld), self.c, self.const,
                                      _p(q), _p(lp), _p(g), _p(self._work))
        return float(lp[0]), g

    def logp_grad_batch(self, Q, n_threads=0):
        """Q float64[C, D] chain-major -> (logp[C], grad[C, D]); one chain per thread."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        C = Q.shape[0]
        lp = np.empty(C)
        g = np.empty_like(Q)
        rc = lib().bgh_oracle_logp_grad_batch(self.M, self.G, _p(self.chi2), _p(self.ld), _p(self.cat, _ip),
                                              self.c, self.const, self.fix, C, _p(Q), _p(lp), _p(g),
                                              int(n_threads))
        if rc != 0:
            raise MemoryError("oracle batch failed rc=%d" % rc)
        return lp, g


def num_threads() -> int:
    return int(lib().bgh_oracle_num_threads())
