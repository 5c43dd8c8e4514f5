"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  The per-pair loop body of kitti_E.cpp:54-201 with the
BFMatcher association of kitti_ba.cpp:641-693, composed from oracle.oracle."""
from __future__ import annotations

import numpy as np

from . import oracle as O


def pair_pipeline(kp0, desc0, kp1, desc1, K, method=O.RANSAC, prob=0.99, thr=1.0, max_iters=1000,
                  norm=O.NORM_HAMMING2, cross_check=True, dist_thresh=50.0, min_trace=2.7,
                  fallback_t=(0.1, 0.1, -0.9), min_t_norm=1e-5, lm_points=48, lambda0=1e-2, eps=1e-8,
                  lm_iters=30, huber_delta=1e-5, lm_revert=1e-9):
    qi, ti, d = O.bf_match(desc0, desc1, norm, cross_check)                 # kitti_ba.cpp:641
    out = points_pipeline(kp0[qi], kp1[ti], K, method, prob, thr, max_iters, dist_thresh, min_trace, fallback_t,   # kitti_ba.cpp:684-693
                          min_t_norm, lm_points, lambda0, eps, lm_iters, huber_delta, lm_revert)
    out["matches"] = (qi, ti, d)
    return out


def points_pipeline(p0, p1, K, method=O.RANSAC, prob=0.99, thr=1.0, max_iters=1000, dist_thresh=50.0, min_trace=2.7,
                    fallback_t=(0.1, 0.1, -0.9), min_t_norm=1e-5, lm_points=48, lambda0=1e-2, eps=1e-8,
                    lm_iters=30, huber_delta=1e-5, lm_revert=1e-9):
    """The loop body after the association, kitti_E.cpp:96-201, for correspondences p0 -> p1 (pixels): whatever made
    them -- descriptor matches (kitti_ba.cpp:684-693) or LK tracks with status 1 (kitti_E.cpp:86-95)."""
    out = {}
    p0 = np.asarray(p0, dtype=np.float32).reshape(-1, 2)
    p1 = np.asarray(p1, dtype=np.float32).reshape(-1, 2)
    Kf = np.asarray(K, dtype=np.float32)
    E, mask, info = O.find_essential_mat(p0, p1, Kf, method, prob, thr, max_iters)   # kitti_E.cpp:98
    out["E"], out["e_mask"], out["e_info"] = E, mask, info
    T0 = np.eye(4)
    ok = E is not None and np.asarray(E).shape == (3, 3)
    n_good = 0
    if ok:
        m = mask == 1
        c0, c1 = p0[m], p1[m]                                               # kitti_E.cpp:106-112
        n_good, R, t, pmask = O.recover_pose(E, c0, c1, Kf, dist_thresh)    # kitti_E.cpp:120
        out["R"], out["t"], out["pose_mask"] = R, t, pmask
    else:
        c0 = c1 = np.zeros((0, 2), np.float32)
        R, t = np.eye(3), np.zeros(3)
    out["n_good"] = n_good
    if (not ok) or np.trace(R) < min_trace:                                 # kitti_E.cpp:128-131
        R, t = np.eye(3), np.array(fallback_t, dtype=np.float64)
    if np.linalg.norm(t) < min_t_norm:                                      # kitti_E.cpp:133-135
        t = np.array(fallback_t, dtype=np.float64)
    T0[:3, :3], T0[:3, 3] = R, t
    out["T0"] = T0.copy()
    out["T"] = T0.copy()
    out["lm_ran"] = False
    out["lm_reverted"] = False
    N = lm_points
    if ok and len(c0) >= N and n_good >= N:                                 # kitti_E.cpp:170-194
        x0 = O.normalize_points(c0[:N], Kf)
        x1 = O.normalize_points(c1[:N], Kf)
        pr = np.concatenate([x0, np.ones((N, 1))], axis=1)[None]
        p_r = np.concatenate([x1, np.ones((N, 1))], axis=1)[None]
        Tout, lm = O.levenberg_marquardt(1, eps, [(0, 0)], [1.0], lambda0, T0[None], pr, p_r, huber_delta, lm_iters)
        out["lm"] = lm
        out["lm_ran"] = True
        if not (lm["r_norm"] <= lm_revert):                                 # kitti_E.cpp:198-200
            out["lm_reverted"] = True
        else:
            out["T"] = Tout[0]
        out["T_lm"] = Tout[0]
    return out


def chain_and_cloud(T_pairs, inliers0, inliers1, K, scales=None):
    """CPU restatement of the reference's post-LM loop body, kitti_E.cpp:203-254, run over a sequence.

    T_pairs[i]: refined 4x4 point transform of pair i; inliers0/1[i]: the E-inlier pixel
    coordinates (cpt0, cpt1_; kitti_E.cpp:106-112) of pair i; scales[i]: ground-truth step length
    (kitti_E.cpp:220).  Returns (all_T (n+1,4,4) incl. the final pose, X (m,3), limits (n,))."""
    n = len(T_pairs)
    scales = np.ones(n) if scales is None else np.asarray(scales, dtype=np.float64)
    Kf = np.asarray(K, dtype=np.float32)
    cT = np.eye(4)
    all_T, X, limits = [], [], []
    for i in range(n):
        T = np.asarray(T_pairs[i], dtype=np.float64)
        dT = np.eye(4)
        t = T[:3, 3] / np.linalg.norm(T[:3, 3])                              # :221
        dT[:3, 3] = t * scales[i]                                            # :222
        dT[:3, :3] = T[:3, :3]                                               # :223
        pT_ = cT.copy()                                                      # :225
        all_T.append(cT.copy())                                              # :227
        cT = cT @ np.linalg.inv(dT)                                          # :228
        R, tt = dT[:3, :3], dT[:3, 3]                                        # :232-233
        limits.append(len(X))                                                # :238
        c0 = O.normalize_points(inliers0[i], Kf)                             # cam_ * (u, v, 1): :240-243
        c1 = O.normalize_points(inliers1[i], Kf)
        for j in range(len(c0)):
            cp = np.array([c0[j, 0], c0[j, 1], 1.0])
            P = np.array([[1.0, 0.0, -c1[j, 0]], [0.0, 1.0, -c1[j, 1]]])     # :245
            A = P @ tt
            B = P @ R @ cp
            if np.linalg.norm(B) > 1e-2:                                     # :248
                d = np.linalg.norm(A) / np.linalg.norm(B)
                X.append(pT_[:3, :3] @ (d * cp) + pT_[:3, 3])                # :250-252
    all_T.append(cT.copy())
    return np.array(all_T), np.array(X).reshape(-1, 3), np.array(limits, dtype=np.int64)
