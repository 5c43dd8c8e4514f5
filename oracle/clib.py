"""CPU ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.  ctypes wrapper of oracle/lm_c.c."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liboracle_lm.so")


class _LmRes(C.Structure):
    _fields_ = [("H_norm", C.c_double), ("r_norm", C.c_double), ("lambda_", C.c_double)]


def build() -> str:
    srcs = [os.path.join(HERE, f) for f in ("lm_c.c", "poly_c.c")]
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(s) for s in srcs):
        subprocess.run(["make", "-C", HERE, "-s"], check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_lm_trace.restype = C.c_int
        _lib.oracle_lm_trace.argtypes = [C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int,
                                         C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(_LmRes),
                                         C.c_void_p]
        _lib.oracle_solve_poly.restype = C.c_int
        _lib.oracle_solve_poly.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    return _lib


def solve_poly(coeffs_ascending, max_iters: int = 300) -> np.ndarray:
    """cv::solvePoly(coeffs, roots, maxIters) restated (oracle/poly_c.c): complex roots in OpenCV's order."""
    c = np.ascontiguousarray(coeffs_ascending, dtype=np.float64)
    n0 = c.shape[0] - 1
    out = np.zeros((n0, 2))
    lib().oracle_solve_poly(c.ctypes.data, n0, int(max_iters), out.ctypes.data)
    return out[:, 0] + 1j * out[:, 1]


def levenberg_marquardt(n_zeta, epsilon, reps, wreps, lambda0, T0s, pr, p_r, huber_delta=1e-5, max_iters=30):
    reps = np.ascontiguousarray(reps, dtype=np.int32).reshape(-1, 2)
    w = np.ascontiguousarray(wreps, dtype=np.float64)
    T = np.array(T0s, dtype=np.float64).reshape(n_zeta, 16).copy()
    pr = np.ascontiguousarray(pr, dtype=np.float64)
    p_r = np.ascontiguousarray(p_r, dtype=np.float64)
    res = _LmRes()
    tr = np.full((int(max_iters), 2), np.nan)
    it = lib().oracle_lm_trace(int(n_zeta), float(epsilon), reps.ctypes.data, w.ctypes.data, reps.shape[0],
                               float(lambda0), int(max_iters), float(huber_delta), T.ctypes.data, pr.ctypes.data,
                               p_r.ctypes.data, int(pr.shape[1]), C.byref(res), tr.ctypes.data)
    trace = [(float(d), None if e != e else float(e)) for d, e in tr[:it]]
    return T.reshape(n_zeta, 4, 4), {"H_norm": res.H_norm, "r_norm": res.r_norm, "lambda": res.lambda_, "iters": it,
                                      "trace": trace}
