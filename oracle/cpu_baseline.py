"""ORACLE (test infrastructure and CPU baseline, not product code).

ctypes wrapper of oracle/_build/libosqpref.so (oracle/osqp_ref.c): the reference's
CasADi->OSQP per-tick solve (reference src/mpc.py:242-258) restated in C, fp64, run on
all host threads.  Used by tests (golden replay) and by bench.py's cpu_baseline /
`--impl reference` legs only.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "_build", "libosqpref.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.run(["make", "-s", "-C", _HERE], check=True)
        L = C.CDLL(LIB)
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p
        L.osqpref_sym_create.restype = vp
        L.osqpref_sym_create.argtypes = [C.c_int]
        L.osqpref_sym_free.argtypes = [vp]
        L.osqpref_sym_nnzL.argtypes = [vp]
        L.osqpref_sym_nvars.argtypes = [vp]
        L.osqpref_work_create.restype = vp
        L.osqpref_work_create.argtypes = [vp]
        L.osqpref_work_free.argtypes = [vp]
        L.osqpref_set_rho.argtypes = [vp, C.c_double]
        L.osqpref_get_rho.restype = C.c_double
        L.osqpref_get_rho.argtypes = [vp]
        L.osqpref_set_tolerances.argtypes = [vp, C.c_double, C.c_double, C.c_int]
        L.osqpref_last_iters.argtypes = [vp]
        L.osqpref_solve.argtypes = [vp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, dp, dp]
        L.osqpref_solve_batch.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, C.c_double,
                                          C.c_double, dp, ip, ip, C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OSQPRefC:
    """One persistent OSQP workspace (rho survives across solves), as CasADi keeps it."""

    def __init__(self, N):
        L = lib()
        self.N = N
        self.sym = L.osqpref_sym_create(N)
        self.work = L.osqpref_work_create(self.sym)
        self.n = L.osqpref_sym_nvars(self.sym)

    def solve(self, x0, r, swing, x_des, mu, delta, g, x_warm=None):
        L = lib()
        c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
        x0, r, swing, x_des = c(x0), c(r), c(swing), c(x_des)
        sol = np.zeros(self.n)
        xw = None if x_warm is None else c(x_warm)
        st = L.osqpref_solve(self.work, _dp(x0), _dp(r), _dp(swing), _dp(x_des), float(mu),
                             float(delta), float(g), None if xw is None else _dp(xw), _dp(sol))
        return sol, st, L.osqpref_last_iters(self.work), L.osqpref_get_rho(self.work)

    def close(self):
        L = lib()
        if self.work:
            L.osqpref_work_free(self.work)
            L.osqpref_sym_free(self.sym)
            self.work = self.sym = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def solve_batch(pb, threads=0, g=-9.81, delta=0.01):
    """Cold-start OSQP-path solve of every problem of a ProblemBatch on `threads` host
    threads (0 = all).  Returns dict(U, iters, status, seconds, threads, kind)."""
    L = lib()
    B, N = pb.B, pb.N
    c = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    x0, r, st, xd, mu = c(pb.x0), c(pb.r), c(pb.stance), c(pb.x_des), c(pb.mu)
    U = np.zeros((B, N, 12))
    iters = np.zeros(B, dtype=np.int32)
    status = np.zeros(B, dtype=np.int32)
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int))
    t0 = time.perf_counter()
    used = L.osqpref_solve_batch(N, B, _dp(x0), _dp(r), _dp(st), _dp(xd), _dp(mu), delta, g,
                                 _dp(U), ip(iters), ip(status), int(threads))
    dt = time.perf_counter() - t0
    return dict(U=U, iters=iters, status=status, seconds=dt, threads=int(used), kind="port",
                mean_iters=float(iters.mean()))
