"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Sequential restatement of the reference's window loops,
`bundle_adjustment` (kitti_ba.cpp:757-905) and `bundle_adjustment_stereo` (kitti_ba.cpp:908-1068),
one `Levenberg_Marquardt` call per window in the reference's order.  The thread polling
(kitti_ba.cpp:793-797) has no CPU counterpart: every reprojection is already in `reprojs`."""
from __future__ import annotations

import numpy as np

from . import oracle as O

MIN_PT = 32


def bundle_adjustment(reprojs, window, stride, num_frames, K, stereo=False, huber_delta=1e-5, lm=None):
    lm_fn = lm or O.levenberg_marquardt
    nodes = 2 * num_frames if stereo else num_frames
    mul = 2 if stereo else 1
    if stereo:                                                       # kitti_ba.cpp:934-941
        win = []
        for i0, i1 in window:
            win += [(2 * i0, 2 * i1), (2 * i0 + 1, 2 * i1), (2 * i0, 2 * i0 + 1)]
    else:
        win = list(window)
    opt_T = [np.eye(4) for _ in range(nodes)]
    optimized = [False] * nodes
    Kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    lms, reverted, starts = [], [], []
    for i in range(0, num_frames, stride):                           # :780 / :943
        w0, w1, stop = nodes + 1, -1, False
        for a, b in win:
            i0, i1 = mul * i + a, mul * i + b
            w0, w1 = min(w0, i0, i1), max(w1, i0, i1)
            if max(i0, i1) >= nodes:                                 # :790-793
                stop = True
                break
        if stop:
            break
        reps, pr, p_r, wreps = [], [], [], []
        for a, b in win:
            i0, i1 = mul * i + a, mul * i + b
            reps.append((a, b - 1) if i1 > i0 else (a - 1, b))       # :810-815
            r = reprojs[(i0, i1)]
            if len(r.p0) < MIN_PT:                                   # :819-824
                wreps.append(0.0)
                pr.append(np.ones((MIN_PT, 3)))
                p_r.append(np.ones((MIN_PT, 3)))
            else:
                wreps.append(r.w if stereo else 1.0)                 # :829-835 / :1000-1008
                h0 = np.column_stack([np.asarray(r.p0[:MIN_PT], dtype=np.float64), np.ones(MIN_PT)])
                h1 = np.column_stack([np.asarray(r.p1[:MIN_PT], dtype=np.float64), np.ones(MIN_PT)])
                pr.append(h0 @ Kinv.T)                               # :838-845
                p_r.append(h1 @ Kinv.T)
        scale = 1.0
        if not stereo and optimized[w0]:                             # :853-856
            scale = np.linalg.norm(opt_T[w0][:3, 3])
        T0s = []
        for j in range(w0, w1):                                      # :857-870
            r = reprojs[(j, j + 1)]
            T = np.eye(4)
            T[:3, :3] = r.R
            T[:3, 3] = np.asarray(r.t).reshape(3)
            T0s.append(T)
        bT0s = [T.copy() for T in T0s]
        Tout, lm = lm_fn(w1 - w0, 1e-8, reps, wreps, 1e-2, T0s, pr, p_r, huber_delta=huber_delta)   # :881
        rev = bool(lm["r_norm"] > 1e-2)                              # :889-891
        Ts = bT0s if rev else list(Tout)
        for j in range(w0, w1):                                      # :895-902
            opt_T[j] = np.array(Ts[j - w0], dtype=np.float64)
            if not stereo:
                opt_T[j][:3, 3] /= scale
            optimized[j] = True
        lms.append([lm["H_norm"], lm["r_norm"], lm["lambda"]])
        reverted.append(rev)
        starts.append(i)
    return np.array(opt_T), np.array(lms).reshape(-1, 3), np.array(reverted, dtype=bool), starts


def match_kp(kps, descs, window, stride, K, counts=None):
    """CPU restatement of `match_kp` (kitti_ba.cpp:583-755), pair by pair in the reference's order: cross-check
    HAMMING2 matching (:602,641), >= 8 matches -> findEssentialMat(LMEDS, 0.99, 0.1) (:702), mask == 1
    compaction (:705-710), recoverPose (:715), rec_mask == 255 compaction (:729-735); else identity and
    (0.1, 0.1, -0.9) (:741-744).  Returns {(i0, i1): (p0, p1, R, t)}.  Checker only."""
    from . import oracle as O
    F = kps.shape[0]
    out = {}
    for i in range(0, F, stride):
        for a, b in window:
            i0, i1 = i + a, i + b
            if (i0, i1) in out:
                continue
            if max(i0, i1) >= F:
                break
            n0 = kps.shape[1] if counts is None else int(counts[i0])
            n1 = kps.shape[1] if counts is None else int(counts[i1])
            qi, ti, _ = O.bf_match(descs[i0][:n0], descs[i1][:n1], O.NORM_HAMMING2, True)
            c0, c1 = kps[i0][qi], kps[i1][ti]
            R, t = np.eye(3), np.array([0.1, 0.1, -0.9])
            p0 = p1 = np.zeros((0, 2), np.float32)
            if len(qi) >= 8:
                E, mask, _ = O.find_essential_mat(c0, c1, np.asarray(K, dtype=np.float32), O.LMEDS, 0.99, 0.1)
                if E is None or np.asarray(E).shape != (3, 3):       # no model: the reference's recoverPose would throw
                    out[(i0, i1)] = (p0, p1, R, t)
                    continue
                c0, c1 = c0[mask == 1], c1[mask == 1]
                n_good, R, t, pmask = O.recover_pose(E, c0, c1, np.asarray(K, dtype=np.float32))
                p0, p1 = c0[pmask == 255], c1[pmask == 255]
            out[(i0, i1)] = (p0, p1, R, np.asarray(t).reshape(3))
    return out
