"""N3: windowed bundle-adjustment orchestration (`bundle_adjustment`, kitti_ba.cpp:757-905, and
`bundle_adjustment_stereo`, kitti_ba.cpp:908-1068) as ONE batched device launch.

The reference walks the windows sequentially, polling a shared map with 20 ms sleeps until the
association thread has produced each reprojection (kitti_ba.cpp:793-797, 1118-1167).  Looking at
the data flow, the Levenberg-Marquardt problem of a window depends only on that window's
reprojections -- its initial chain is rebuilt from `reprojs[(j, j+1)]` every time
(kitti_ba.cpp:856-859: `if(!optimized[j] || true)`) -- so all windows are independent LM
problems of identical shape.  They go to the GPU together (`epivo_lm_rt_batch`, one CTA per
window); only the cheap tail is sequential and stays on the host, exactly as written:
revert if `r_norm > 1e-2` (kitti_ba.cpp:889-891), divide the translations by the scale carried
from the previous window (mono only, kitti_ba.cpp:853-856, 898-901), later windows overwrite
the overlap.  Windows shard across GPUs like frame pairs do (contiguous ranges of window starts).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import api

MIN_PT = 32          # kitti_ba.cpp:777, 927
LM_EPS = 1e-8        # kitti_ba.cpp:881
LM_LAMBDA0 = 1e-2
REVERT_R_NORM = 1e-2  # kitti_ba.cpp:889


@dataclass
class Reproj:
    """`struct reproj` (kitti_ba.cpp:167-175): matched pixels of a frame pair + its recoverPose estimate."""
    p0: np.ndarray                      # (n, 2) pixels in frame i0
    p1: np.ndarray                      # (n, 2) pixels in frame i1
    R: np.ndarray = field(default_factory=lambda: np.eye(3))
    t: np.ndarray = field(default_factory=lambda: np.zeros(3))
    w: float = 1.0                      # reprojection weight; 0 freezes R, t (stereo extrinsics)


def expand_stereo_window(window):
    """kitti_ba.cpp:934-941: frame f -> nodes 2f (left), 2f+1 (right)."""
    out = []
    for i0, i1 in window:
        out += [(2 * i0, 2 * i1), (2 * i0 + 1, 2 * i1), (2 * i0, 2 * i0 + 1)]
    return out


def window_reps(window):
    """kitti_ba.cpp:810-815: window entry (first, second) -> zeta span of its reprojection."""
    reps = []
    for a, b in window:
        assert a != b
        reps.append((a, b - 1) if b > a else (a - 1, b))
    return reps


def window_starts(window, stride, num_nodes, node_step=1):
    """Window start indices the reference loop visits before its `_exit` break (kitti_ba.cpp:780-801)."""
    span = max(max(a, b) for a, b in window)
    starts = []
    for i in range(0, num_nodes // node_step, stride):
        if node_step * i + span >= num_nodes:
            break
        starts.append(i)
    return starts


def normalize(px, Kinv):
    """cam_ * (u, v, 1) (kitti_ba.cpp:838-845), float64."""
    px = np.asarray(px, dtype=np.float64)
    h = np.column_stack([px[:, 0], px[:, 1], np.ones(len(px))])
    return h @ Kinv.T


def assemble(reprojs, window, starts, K, stereo=False):
    """Per-window LM inputs, stacked: T0s (B, nz, 4, 4), pr / p_r (B, n_rep, 32, 3), wreps (B, n_rep)."""
    Kinv = np.linalg.inv(np.asarray(K, dtype=np.float64))
    step = 2 if stereo else 1
    lo = min(min(a, b) for a, b in window)
    hi = max(max(a, b) for a, b in window)
    nz, B, n_rep = hi - lo, len(starts), len(window)
    T0s = np.tile(np.eye(4), (B, nz, 1, 1))
    pr = np.ones((B, n_rep, MIN_PT, 3))
    p_r = np.ones((B, n_rep, MIN_PT, 3))
    w = np.zeros((B, n_rep))
    for b, i in enumerate(starts):
        base = step * i
        for j, (a, c) in enumerate(window):
            r = reprojs[(base + a, base + c)]
            if len(r.p0) < MIN_PT:                      # "Bad pts": dummy ones, weight 0 (kitti_ba.cpp:819-824)
                continue
            w[b, j] = r.w if stereo else 1.0
            pr[b, j] = normalize(r.p0[:MIN_PT], Kinv)
            p_r[b, j] = normalize(r.p1[:MIN_PT], Kinv)
        for k in range(nz):                             # kitti_ba.cpp:856-868: chain from the pairwise estimates
            r = reprojs[(base + lo + k, base + lo + k + 1)]
            T0s[b, k, :3, :3] = r.R
            T0s[b, k, :3, 3] = np.asarray(r.t).reshape(3)
    return T0s, pr, p_r, w, nz, lo


def bundle_adjustment(reprojs, window, stride, num_frames, K, stereo=False, huber_delta=1e-5, ctx=None):
    """All windows in one device launch.  Returns (opt_T (nodes, 4, 4), lm (B, 3) [H_norm, r_norm, lambda],
    reverted (B,) bool, starts).  nodes = num_frames (mono) or 2 * num_frames (stereo)."""
    win = expand_stereo_window(window) if stereo else list(window)
    nodes = 2 * num_frames if stereo else num_frames
    step = 2 if stereo else 1
    starts = window_starts(win, stride, nodes, step)
    opt_T = np.tile(np.eye(4), (nodes, 1, 1))
    if not starts:
        return opt_T, np.zeros((0, 3)), np.zeros(0, bool), starts
    T0s, pr, p_r, w, nz, lo = assemble(reprojs, win, starts, K, stereo)
    reps = window_reps(win)
    T_opt, lm, _ = api.Levenberg_Marquardt_batch(nz, LM_EPS, reps, w, LM_LAMBDA0, T0s, pr, p_r,
                                                 huber_delta=huber_delta, ctx=ctx)
    return (*finish(opt_T, T0s, T_opt, lm, starts, nz, lo, step, stereo), starts)


def finish(opt_T, T0s, T_opt, lm, starts, nz, lo, step, stereo):
    """The sequential tail of the reference loop (kitti_ba.cpp:853-856, 889-903 / 1054-1066)."""
    optimized = np.zeros(len(opt_T), dtype=bool)
    reverted = np.zeros(len(starts), dtype=bool)
    for b, i in enumerate(starts):
        w0 = step * i + lo
        scale = 1.0
        if not stereo and optimized[w0]:
            scale = np.linalg.norm(opt_T[w0][:3, 3])
        Ts = T_opt[b]
        if lm[b, 1] > REVERT_R_NORM:
            Ts = T0s[b]
            reverted[b] = True
        for k in range(nz):
            opt_T[w0 + k] = Ts[k]
            if not stereo:
                opt_T[w0 + k][:3, 3] /= scale
            optimized[w0 + k] = True
    return opt_T, np.asarray(lm), reverted


def window_pairs(window, stride, num_frames):
    """The (i0, i1) keys `match_kp` visits, in its order, without the repeats it skips (kitti_ba.cpp:603-615)."""
    pairs, seen = [], set()
    for i in range(0, num_frames, stride):
        for a, b in window:
            i0, i1 = i + a, i + b
            if (i0, i1) in seen:
                continue
            if max(i0, i1) >= num_frames:
                break
            seen.add((i0, i1))
            pairs.append((i0, i1))
    return pairs


def match_kp(kps, descs, window, stride, K, counts=None, ctx=None):
    """`match_kp` (kitti_ba.cpp:583-755) for a whole drive in ONE pipeline pass: every window pair
    (i + first, i + second) is matched (BFMatcher NORM_HAMMING2 cross-check, :602,641), filtered by
    findEssentialMat(LMEDS, 0.99, 0.1) (:702) and recoverPose (:715), and the `reprojs` map the bundle
    adjustment consumes is returned: p0 / p1 = matched pixels that are E-inliers (mask == 1) AND pass the
    cheirality test (rec_mask == 255), R / t = recoverPose's; fewer than 8 matches -> identity and
    (0.1, 0.1, -0.9) with no points (:741-744).

    kps (F, kp, 2) f32, descs (F, kp, 32) u8, counts (F,) optional keypoints per frame.  The reference's
    association thread does this pair by pair on the CPU while the BA thread polls the map; here the explicit
    pair list (`epivo_seq_set_pairs`) makes it a single batch."""
    kps = np.ascontiguousarray(kps, dtype=np.float32)
    descs = np.ascontiguousarray(descs, dtype=np.uint8)
    F, kp = kps.shape[0], kps.shape[1]
    pairs = window_pairs(window, stride, F)
    reprojs = {}
    if not pairs:
        return reprojs
    pipe = api.SequencePipeline(F, kp, ctx=ctx, max_pairs=max(len(pairs), F - 1))
    try:
        if counts is not None:
            pipe.set_counts(counts)
        pipe.set_pairs([p[0] for p in pairs], [p[1] for p in pairs])
        prm = api.default_params(np.asarray(K, dtype=np.float32), method=api.LMEDS, prob=0.99, threshold=0.1)
        res = pipe.process(prm, kps, descs)
        for p, (i0, i1) in enumerate(pairs):
            r = Reproj(p0=np.zeros((0, 2), np.float32), p1=np.zeros((0, 2), np.float32), R=np.eye(3),
                       t=np.array([0.1, 0.1, -0.9]))
            # no model (findEssentialMat would return an empty Mat): keep the identity / (0.1, 0.1, -0.9) start
            if res[p]["n_matches"] >= 8 and res[p]["n_inliers"] > 0:
                qi, ti, _ = pipe.matches(p)
                em, pm = pipe.masks(p)
                keep = np.flatnonzero(em == 1)[pm == 255]
                r.p0 = kps[i0][qi[keep]]
                r.p1 = kps[i1][ti[keep]]
                r.R = res[p]["R"].copy()
                r.t = res[p]["t"].copy()
            reprojs[(i0, i1)] = r
    finally:
        pipe.close()
    return reprojs
