"""CPU ORACLE -- TEST / BASELINE INFRASTRUCTURE ONLY.

ctypes wrapper of oracle/_ref/libref_lm*.so: the reference's OWN LM translation unit
(/root/reference/jac_Rt_gen_.cpp:23-478, plus sequence.hpp's generators), compiled unmodified against
the Eigen/Sophus stand-ins in oracle/ref_shim/ (recipe: oracle/Makefile, glue: oracle/ref_lm_tu.cpp).

    REF     libref_lm.so      unmodified; huber_delta = 1e-5 (jac_Rt_gen_.cpp:17)
    REF_D1  libref_lm_d1.so   PATCHED in the compile pipe: that one constant set to 1.0
                              (the value test_jac_Rt_gen.cpp:16 tests with)
    DEMO    libref_demo.so    test_jac_Rt_gen.cpp unmodified: res / Dr_Deps / forward RepJacobian at
                              huber_delta = 1.0 and the seeded convergence demo (main renamed)

The libraries are built where /root/reference exists (the build container) and travel to the GPU box
as files; `available()` is False where neither the files nor the sources exist.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REFERENCE_SRC = os.environ.get("EPIVO_REFERENCE", "/root/reference")
_FILES = {"ref_": "libref_lm.so", "refd1_": "libref_lm_d1.so", "refdemo_": "libref_demo.so"}


def build() -> bool:
    """(Re)build oracle/_ref when the reference sources are present; returns availability."""
    if os.path.exists(os.path.join(REFERENCE_SRC, "jac_Rt_gen_.cpp")):
        subprocess.run(["make", "-C", HERE, "-s", f"REF={REFERENCE_SRC}"], check=True)
    return available()


def available() -> bool:
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in _FILES.values())


_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class RefLib:
    def __init__(self, prefix: str):
        self.prefix = prefix
        path = os.path.join(REF_DIR, _FILES[prefix])
        if not os.path.exists(path):
            build()
        self.lib = C.CDLL(path)
        f = self._fn
        f("huber_delta", C.c_double, [])
        f("res", C.c_int, [_dp, _dp, _dp, _dp, C.c_int, _dp])
        f("dr_deps", C.c_int, [_dp, _dp, _dp, _dp, C.c_int, C.c_int, _dp])
        f("rep_jacobian", C.c_int, [C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp])
        f("se3_exp", None, [_dp, _dp])
        f("gen_scene_sequence", C.c_int, [C.c_uint, C.c_int, C.c_int, _ip, C.c_int, _dp, _dp, _dp, _dp, _dp])
        if prefix == "refdemo_":
            f("demo", C.c_int, [C.c_uint, C.c_char_p, C.c_int])
        else:
            f("lm", C.c_int, [C.c_int, C.c_double, _ip, _dp, C.c_int, C.c_double, _dp, _dp, _dp, C.c_int, _dp,
                              C.POINTER(C.c_int), _dp, C.c_int, C.POINTER(C.c_int)])

    def _fn(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype, fn.argtypes = restype, argtypes
        setattr(self, "_" + name, fn)

    @property
    def huber_delta(self) -> float:
        return self._huber_delta()

    def res(self, R0, t0, p, p_):
        p, p_ = _c(p), _c(p_)
        r = np.zeros(p.shape[0])
        self._res(_c(R0), _c(t0), p, p_, p.shape[0], r)
        return r

    def dr_deps(self, Tl0, Tr0, p, p_, reverse: bool):
        p, p_ = _c(p), _c(p_)
        J = np.zeros((p.shape[0], 6))
        rc = self._dr_deps(_c(Tl0), _c(Tr0), p, p_, p.shape[0], int(bool(reverse)), J)
        if rc:
            raise ValueError("this reference build has no `reverse` argument")
        return J

    def rep_jacobian(self, T0s, zeta, src, tgt, p, p_):
        T = _c(np.asarray(T0s, dtype=np.float64).reshape(-1, 16))
        p, p_ = _c(p), _c(p_)
        J = np.zeros((p.shape[0], 6))
        self._rep_jacobian(T.shape[0], T, zeta, src, tgt, p, p_, p.shape[0], J)
        return J

    def se3_exp(self, a):
        T = np.zeros((4, 4))
        self._se3_exp(_c(a), T)
        return T

    def gen_scene_sequence(self, seed, N, n_zeta, reps):
        """sequence.hpp:106-159 with srand(seed): (Ts, T0s, Xr, pr, p_r)."""
        rp = np.ascontiguousarray(reps, dtype=np.int32).reshape(-1, 2)
        n_rep = rp.shape[0]
        Ts, T0s = np.zeros((n_zeta, 4, 4)), np.zeros((n_zeta, 4, 4))
        Xr, pr, p_r = (np.zeros((n_rep, N, 3)) for _ in range(3))
        self._gen_scene_sequence(int(seed), N, n_zeta, rp, n_rep, Ts, T0s, Xr, pr, p_r)
        return Ts, T0s, Xr, pr, p_r

    def levenberg_marquardt(self, n_zeta, epsilon, reps, wreps, lambda0, T0s, pr, p_r):
        """The 9-argument Levenberg_Marquardt (jac_Rt_gen_.cpp:287-296); 30 iterations, this build's delta.

        Returns (T0s, info): info has H_norm / r_norm / lambda (the reference's LM_res), `nan_break`
        and `iters` = completed iterations, recovered from lambda = lambda0 * 5^rejects / 2^accepts
        (every completed iteration does exactly one of the two, :456-467; the reference returns no count)."""
        rp = np.ascontiguousarray(reps, dtype=np.int32).reshape(-1, 2)
        w = np.ascontiguousarray(wreps, dtype=np.float64)
        T = np.array(T0s, dtype=np.float64).reshape(n_zeta, 16).copy()
        pr, p_r = _c(pr), _c(p_r)
        out = np.zeros(3)
        nanb, ntr = C.c_int(0), C.c_int(0)
        tr = np.zeros((80, 2))
        self._lm(int(n_zeta), float(epsilon), rp, w, rp.shape[0], float(lambda0), T, pr, p_r, int(pr.shape[1]), out,
                 C.byref(nanb), tr, tr.shape[0], C.byref(ntr))
        acc, rej = lambda_steps(out[2], lambda0)
        return T.reshape(n_zeta, 4, 4), {"H_norm": out[0], "r_norm": out[1], "lambda": out[2],
                                          "nan_break": bool(nanb.value), "accepts": acc, "rejects": rej,
                                          "iters": acc + rej,
                                          "trace": parse_trace(tr[:ntr.value], 6 * n_zeta, rp.shape[0] * pr.shape[1])}

    def demo(self, seed: int, cwd: str | None = None) -> str:
        """Runs test_jac_Rt_gen.cpp's main() with srand(seed); it writes est.pose / gt.pose into `cwd`."""
        import tempfile
        buf = C.create_string_buffer(1 << 16)
        old = os.getcwd()
        with tempfile.TemporaryDirectory() as tmp:
            os.chdir(cwd or tmp)
            try:
                self._demo(int(seed), buf, len(buf))
            finally:
                os.chdir(old)
        return buf.value.decode()


def parse_trace(tr, D, rep_N):
    """(size, value) norm() records of one LM call -> list of (|delta|, curr_E or None) per started iteration.
    The last two records are |H| and |r0| after the loop (jac_Rt_gen_.cpp:473-474)."""
    body = [(int(s), v) for s, v in tr[:-2]]
    its, k = [], 0
    while k < len(body):
        assert body[k][0] == D, body
        dn = body[k][1]
        k += 1
        if k < len(body) and body[k][0] == rep_N and (rep_N != D or True) and not (dn != dn):
            # a |delta| record is followed by the candidate |r0| unless the loop broke on |delta| < epsilon
            if body[k][0] == rep_N and not (rep_N == D and k + 1 == len(body) and False):
                its.append((dn, body[k][1]))
                k += 1
                continue
        its.append((dn, None))
    return its


def lambda_steps(lam: float, lambda0: float, max_iters: int = 64):
    """(accepts, rejects) with lam == lambda0 * 5**rejects / 2**accepts (unique: 2 and 5 are coprime)."""
    best = None
    for rej in range(max_iters + 1):
        a = math.log2(lambda0 * 5.0 ** rej / lam)
        acc = int(round(a))
        if acc < 0 or acc + rej > max_iters:
            continue
        if abs(lambda0 * 5.0 ** rej / 2.0 ** acc - lam) <= 1e-9 * lam:
            if best is not None:
                raise ValueError("ambiguous lambda factorisation")
            best = (acc, rej)
    if best is None:
        raise ValueError(f"lambda {lam} is not lambda0 * 5^b / 2^a")
    return best


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


_cache: dict = {}


def ref() -> RefLib:          # unmodified, delta = 1e-5
    if "ref_" not in _cache:
        _cache["ref_"] = RefLib("ref_")
    return _cache["ref_"]


def ref_d1() -> RefLib:       # patched constant, delta = 1.0
    if "refd1_" not in _cache:
        _cache["refd1_"] = RefLib("refd1_")
    return _cache["refd1_"]


def demo() -> RefLib:         # test_jac_Rt_gen.cpp, delta = 1.0
    if "refdemo_" not in _cache:
        _cache["refdemo_"] = RefLib("refdemo_")
    return _cache["refdemo_"]
