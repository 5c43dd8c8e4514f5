"""CPU BASELINE -- used only by bench.py (`cpu_baseline`, `--impl reference`) and tests.

The reference's per-pair path on host cores: the OpenCV calls the reference itself makes
(through the cv2 wheel: BFMatcher.match, findEssentialMat, recoverPose -- kitti_ba.cpp:641,
kitti_E.cpp:98-120) followed by the reference's OWN Levenberg_Marquardt -- jac_Rt_gen_.cpp compiled
unmodified into oracle/_ref/libref_lm.so (oracle/Makefile; built where /root/reference exists, the file
travels to the GPU box).  Where that library is absent the plain-C restatement (oracle/lm_c.c) runs
instead; if cv2 cannot be imported the numpy restatement (oracle.oracle) is timed.  The JSON line says
which (`cpu_baseline.kind`, `.sample`).
"""
from __future__ import annotations

import os
import time

import numpy as np

try:
    import cv2
    HAVE_CV2 = True
except Exception:                                       # pragma: no cover
    cv2 = None
    HAVE_CV2 = False

from . import clib
from . import oracle as O
from . import reflib

_SEQ = {}


def lm_kind() -> str:
    """"reference" when oracle/_ref/libref_lm.so (the reference's own LM) is loadable, else "port"."""
    return "reference" if reflib.available() else "port"


def describe() -> str:
    if not HAVE_CV2:
        return "numpy restatement (cv2 missing)"
    lm = ("the reference's own Levenberg_Marquardt (jac_Rt_gen_.cpp compiled unmodified, oracle/_ref, g++ -O0 as its "
          "functions without return statements require)" if lm_kind() == "reference"
          else "plain-C restatement of its LM (oracle/_ref absent)")
    return "cv2 %s BFMatcher/findEssentialMat/recoverPose (the OpenCV calls the reference makes) + %s" % (cv2.__version__, lm)


def lm_ms_per_pair(n: int = 5) -> dict:
    """Time of one kitti_E-shaped LM call (1 zeta, 48 points, 30 iterations) in the two CPU implementations, so that
    a reader can see what the choice of LM build does to the baseline: the reference build is g++ -O0 (its functions
    without return statements rule out optimisation) against a stand-in Eigen; the C port is -O2."""
    from epivo_b200 import synth
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(9, 48, 1, [(0, 0)], noise_rot=1e-2, noise_tr=1e-2)
    out = {}
    impls = [("c_port", lambda: clib.levenberg_marquardt(1, 1e-8, [(0, 0)], [1.0], 1e-2, T0s, pr, p_r, 1e-5, 30))]
    if lm_kind() == "reference":
        impls.append(("reference_build", lambda: reflib.ref().levenberg_marquardt(1, 1e-8, [(0, 0)], [1.0], 1e-2, T0s, pr, p_r)))
    for name, fn in impls:
        fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        out[name] = (time.perf_counter() - t0) / n * 1e3
    return out


def _lm(T, pr, p_r, huber_delta):
    """kitti_E.cpp:196 -> (T_out (1,4,4), r_norm)"""
    if lm_kind() == "reference" and huber_delta == 1e-5:
        Tl, lm = reflib.ref().levenberg_marquardt(1, 1e-8, [(0, 0)], [1.0], 1e-2, T, pr, p_r)
    else:
        Tl, lm = clib.levenberg_marquardt(1, 1e-8, [(0, 0)], [1.0], 1e-2, T, pr, p_r, huber_delta, 30)
    return Tl, lm["r_norm"]


def pair_cv2(kp0, d0, kp1, d1, Kf, method=8, prob=0.99, thr=1.0, lm_points=48, huber_delta=1e-5,
             norm=7, ratio=None):
    """One frame pair through the reference's CPU path.  Returns (T 4x4, n_matches, n_inliers, n_good).
    ratio=None: BFMatcher(norm, crossCheck=true).match (kitti_ba.cpp:602,641); else knnMatch(k=2) + Lowe ratio."""
    if ratio is None:
        ms = cv2.BFMatcher(norm, True).match(d0, d1)
    else:
        ms = [a for a, b in cv2.BFMatcher(norm, False).knnMatch(d0, d1, k=2) if a.distance < ratio * b.distance]
    qi = np.fromiter((m.queryIdx for m in ms), dtype=np.int64, count=len(ms))
    ti = np.fromiter((m.trainIdx for m in ms), dtype=np.int64, count=len(ms))
    p0, p1 = kp0[qi], kp1[ti]                                             # kitti_ba.cpp:684-693
    T = np.eye(4)
    if len(p0) < 5:
        T[:3, 3] = (0.1, 0.1, -0.9)
        return T, len(p0), 0, 0
    E, mask = cv2.findEssentialMat(p0, p1, Kf, method, prob, thr)         # kitti_E.cpp:98-104
    if E is None or E.shape != (3, 3):
        T[:3, 3] = (0.1, 0.1, -0.9)
        return T, len(p0), 0, 0
    m = mask.ravel() == 1
    c0, c1 = p0[m], p1[m]                                                 # kitti_E.cpp:106-112
    n_good, R, t, rm = cv2.recoverPose(E, c0, c1, Kf)                     # kitti_E.cpp:120
    t = t.ravel()
    if np.trace(R) < 2.7:                                                 # kitti_E.cpp:128-135
        R, t = np.eye(3), np.array([0.1, 0.1, -0.9])
    if np.linalg.norm(t) < 1e-5:
        t = np.array([0.1, 0.1, -0.9])
    T[:3, :3], T[:3, 3] = R, t
    N = lm_points
    if len(c0) >= N and int((rm.ravel() == 255).sum()) >= N:              # kitti_E.cpp:170-201
        x0 = O.normalize_points(c0[:N], Kf)
        x1 = O.normalize_points(c1[:N], Kf)
        pr = np.concatenate([x0, np.ones((N, 1))], axis=1)[None]
        p_r = np.concatenate([x1, np.ones((N, 1))], axis=1)[None]
        Tl, r_norm = _lm(T[None], pr, p_r, huber_delta)
        if r_norm <= 1e-9:
            T = Tl[0]
    return T, len(p0), int(m.sum()), int(n_good)


def pair_numpy(kp0, d0, kp1, d1, Kf, method=8, prob=0.99, thr=1.0, **kw):
    from . import pipeline
    o = pipeline.pair_pipeline(kp0, d0, kp1, d1, Kf, method, prob, thr)
    return o["T"], len(o["matches"][0]), int(o["e_mask"].sum()), int(o["n_good"])


def _init(kps, descs, Kf, method, prob, thr, norm=7, ratio=None):
    _SEQ["kps"], _SEQ["descs"], _SEQ["K"], _SEQ["args"] = kps, descs, Kf, (method, prob, thr)
    _SEQ["match"] = (norm, ratio)
    if HAVE_CV2:
        cv2.setNumThreads(1)                                # pair-parallel: one core per pair
    clib.lib()
    if lm_kind() == "reference":
        reflib.ref()


def _work(i):
    kps, descs, Kf = _SEQ["kps"], _SEQ["descs"], _SEQ["K"]
    method, prob, thr = _SEQ["args"]
    if HAVE_CV2:
        norm, ratio = _SEQ["match"]
        T, nm, ni, ng = pair_cv2(kps[i], descs[i], kps[i + 1], descs[i + 1], Kf, method, prob, thr, norm=norm, ratio=ratio)
    else:
        T, nm, ni, ng = pair_numpy(kps[i], descs[i], kps[i + 1], descs[i + 1], Kf, method, prob, thr)
    return i, T, nm, ni, ng


# Worker processes are SPAWNED, not forked: the callers (bench.py after its GPU legs, the GPU census test) have a
# live CUDA context and helper threads by then, and a fork()ed child of a multi-threaded process can deadlock on a
# lock some other thread held at the time of the fork (seen once as a 20-minute hang of the census).
_MP = "spawn"
POOL_TIMEOUT_S = 900       # a pool whose workers cannot start would otherwise wait forever


class CpuPool:
    """Pair-parallel CPU workers (how a CPU user would saturate the box): `cores` processes,
    cv2.setNumThreads(1) each.  The sequence is sent to every worker once (initializer), not per task."""

    def __init__(self, kps, descs, K, method=8, prob=0.99, thr=1.0, cores=None, norm=7, ratio=None):
        import multiprocessing as mp
        self.cores = cores or (os.cpu_count() or 1)
        clib.build()
        Kf = np.asarray(K, dtype=np.float32)
        self.pool = mp.get_context(_MP).Pool(self.cores, initializer=_init,
                                                initargs=(kps, descs, Kf, method, prob, thr, norm, ratio))

    def run(self, pair_indices):
        """-> (pairs/s, results)"""
        t0 = time.perf_counter()
        res = self.pool.map_async(_work, list(pair_indices), chunksize=1).get(timeout=POOL_TIMEOUT_S)
        dt = time.perf_counter() - t0
        return len(res) / dt, res

    def close(self):
        self.pool.terminate()
        self.pool.join()


def single_process_rate(kps, descs, K, n_pairs, method=8, prob=0.99, thr=1.0, threads=None, norm=7, ratio=None):
    """SURVEY 8d's second CPU arrangement: ONE process, OpenCV's own thread pool over `threads` cores
    (cv2.setNumThreads), pairs one after the other.  -> pairs/s (one warm-up pair is not timed)."""
    if not HAVE_CV2:
        return None
    threads = threads or (os.cpu_count() or 1)
    Kf = np.asarray(K, dtype=np.float32)
    clib.build()
    cv2.setNumThreads(threads)
    try:
        pair_cv2(kps[0], descs[0], kps[1], descs[1], Kf, method, prob, thr, norm=norm, ratio=ratio)
        t0 = time.perf_counter()
        for i in range(n_pairs):
            pair_cv2(kps[i], descs[i], kps[i + 1], descs[i + 1], Kf, method, prob, thr, norm=norm, ratio=ratio)
        return n_pairs / (time.perf_counter() - t0)
    finally:
        cv2.setNumThreads(1)


_WIN = {}


def _win_init(args):
    _WIN["args"] = args
    clib.lib()


def _win_work(b):
    nz, reps, data, delta = _WIN["args"]
    Ts, T0, pr, p_r = data[b % len(data)]
    clib.levenberg_marquardt(nz, 1e-8, reps, [1.0] * len(reps), 1e-2, T0, pr, p_r, huber_delta=delta, max_iters=30)
    return b


def windows_rate(data, nz, reps, n_windows, cores=None, huber_delta=1.0):
    """kitti_ba windows (config 5) on the host cores: the plain-C restatement of the reference's Levenberg_Marquardt
    (the Eigen / Sophus original cannot be built here), one window per process at a time.  -> windows/s."""
    import multiprocessing as mp
    cores = cores or (os.cpu_count() or 1)
    clib.build()
    clib.lib()
    pool = mp.get_context(_MP).Pool(cores, initializer=_win_init, initargs=((nz, reps, data, huber_delta),))
    try:
        pool.map_async(_win_work, range(cores), chunksize=1).get(timeout=POOL_TIMEOUT_S)   # warm-up: load the library
        t0 = time.perf_counter()
        pool.map_async(_win_work, range(n_windows), chunksize=1).get(timeout=POOL_TIMEOUT_S)
        return n_windows / (time.perf_counter() - t0)
    finally:
        pool.close()
        pool.join()


# ---- live-cv2 parity census (tests/test_gpu_cv2_census.py, bench.py `parity_vs_cv2`) -------------------------------
# The reference's findEssentialMat call shapes (method, prob, threshold), by call site.
CALL_SHAPES = {
    "kitti.cpp:101": (8, 0.99, 1.0),          # RANSAC
    "kitti_E.cpp:101": (4, 0.99, 0.01),       # LMEDS (threshold unused by LMedS)
    "euroc_E.cpp:205": (8, 0.99, 0.3),        # RANSAC, EuRoC camera, 1500 keypoints
    "kitti_ba.cpp:232": (8, 0.95, 0.05),
    "kitti_ba.cpp:308": (8, 0.99, 0.05),
    "kitti_ba.cpp:702": (4, 0.99, 0.1),
}


def _census_work(i):
    kps, descs, Kf = _SEQ["kps"], _SEQ["descs"], _SEQ["K"]
    norm, _ = _SEQ["match"]
    ms = cv2.BFMatcher(norm, True).match(descs[i], descs[i + 1])
    qi = np.fromiter((m.queryIdx for m in ms), dtype=np.int32, count=len(ms))
    ti = np.fromiter((m.trainIdx for m in ms), dtype=np.int32, count=len(ms))
    p0, p1 = kps[i][qi], kps[i + 1][ti]
    out = {"pair": i, "qi": qi, "ti": ti, "shapes": {}}
    for name in _SEQ["shapes"]:
        method, prob, thr = CALL_SHAPES[name]
        E, mask = cv2.findEssentialMat(p0, p1, Kf, method, prob, thr)
        rec = {"E": None}
        if E is not None and E.shape == (3, 3):
            m = mask.ravel() == 1
            n_good, R, t, rm = cv2.recoverPose(E, p0[m], p1[m], Kf)
            rec = {"E": E, "e_mask": mask.ravel().astype(np.uint8), "n_good": int(n_good), "R": R, "t": t.ravel(),
                   "pose_mask": rm.ravel().astype(np.uint8)}
        out["shapes"][name] = rec
    return out


def census(kps, descs, K, shapes, norm=7, cores=None):
    """cv2's matches, E, {0,1} mask, recoverPose count / R / t / {0,255} mask for every consecutive pair of the
    sequence and every named call shape -- the live reference the GPU pipeline is compared with, pair-parallel."""
    import multiprocessing as mp
    assert HAVE_CV2
    cores = cores or (os.cpu_count() or 1)
    Kf = np.asarray(K, dtype=np.float32)
    _SEQ["shapes"] = list(shapes)
    pool = mp.get_context(_MP).Pool(cores, initializer=_census_init, initargs=(kps, descs, Kf, norm, list(shapes)))
    try:
        return pool.map_async(_census_work, range(kps.shape[0] - 1), chunksize=1).get(timeout=POOL_TIMEOUT_S)
    finally:
        pool.close()
        pool.join()


def _census_init(kps, descs, Kf, norm, shapes):
    _SEQ["kps"], _SEQ["descs"], _SEQ["K"], _SEQ["match"], _SEQ["shapes"] = kps, descs, Kf, (norm, None), shapes
    cv2.setNumThreads(1)
