"""Seeded synthetic KITTI-/EuRoC-shaped keypoint + descriptor workloads.

The reference reads real KITTI / EuRoC images (kitti_E.cpp:46-65, euroc_E.cpp:134-172) and
has no synthetic image-level fixtures; BASELINE.json's configs are defined on synthetic
keypoint/descriptor sets.  This module is the generator for those configs.
It only produces INPUTS (numpy arrays) for the benchmark and the tests; it is not on the
product path and contains no geometry solver.

Conventions
-----------
* keypoints: float32 (n, 2) pixel coordinates (x, y) -- `vector<Point2f>` in the reference.
* descriptors: uint8 (n, 32) -- ORB-style 256-bit rows (`kitti_ba.cpp:128`).
* motion (R, t): *point* transform X1 = R X0 + t, the convention `recoverPose` returns.
"""
from __future__ import annotations

import dataclasses

import numpy as np

KITTI_K = np.array([[718.8560, 0.0, 607.1928],
                    [0.0, 718.8560, 185.2157],
                    [0.0, 0.0, 1.0]], dtype=np.float64)      # kitti_E.cpp:38-40
KITTI_SIZE = (1241, 376)
EUROC_K = np.array([[435.2047, 0.0, 367.4517],
                    [0.0, 435.2047, 252.2009],
                    [0.0, 0.0, 1.0]], dtype=np.float64)      # euroc_E.cpp:150-152
EUROC_SIZE = (752, 480)


def seed_for(cfg: int, index: int) -> int:
    """seed = 10_000 * cfg + pair_index."""
    return 10_000 * int(cfg) + int(index)


def rodrigues(w: np.ndarray) -> np.ndarray:
    w = np.asarray(w, dtype=np.float64).reshape(3)
    th = float(np.linalg.norm(w))
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]], dtype=np.float64)
    if th < 1e-12:
        return np.eye(3) + Kx
    return np.eye(3) + (np.sin(th) / th) * Kx + ((1 - np.cos(th)) / th ** 2) * (Kx @ Kx)


def flip_bits(desc: np.ndarray, rng: np.random.Generator) -> np.ndarray:
    """Flip every bit independently with p = 5/64 = 0.078 (inlier Hamming ~ 20 +- 4).

    Built from six random byte planes so that a 4541-frame sequence can be generated in
    seconds: p = P(a&b&c&d) + P(a&b&c&~d&e&f) = 1/16 + 1/64.
    """
    shp = desc.shape
    r = rng.integers(0, 256, size=(6,) + shp, dtype=np.uint8)
    abc = r[0] & r[1] & r[2]
    mask = (abc & r[3]) | (abc & ~r[3] & r[4] & r[5])
    return desc ^ mask


@dataclasses.dataclass
class Pair:
    kp0: np.ndarray        # (n0, 2) f32
    kp1: np.ndarray        # (n1, 2) f32
    desc0: np.ndarray      # (n0, 32) u8
    desc1: np.ndarray      # (n1, 32) u8
    K: np.ndarray          # (3, 3) f64
    R: np.ndarray          # ground-truth rotation (point transform)
    t: np.ndarray          # ground-truth translation
    gt_match: np.ndarray   # (n0,) int32: index into frame 1 of the true correspondence, -1 if none


def _project(K, X):
    x = X[:, 0] / X[:, 2]
    y = X[:, 1] / X[:, 2]
    return np.stack([K[0, 0] * x + K[0, 2], K[1, 1] * y + K[1, 2]], axis=1)


def _backproject(K, px, depth):
    x = (px[:, 0] - K[0, 2]) / K[0, 0]
    y = (px[:, 1] - K[1, 2]) / K[1, 1]
    return np.stack([x * depth, y * depth, depth], axis=1)


def make_pair(seed: int, n: int = 2000, K: np.ndarray = KITTI_K, size=KITTI_SIZE,
              depth=(5.0, 60.0), rot_sigma: float = 0.01, t_mean=(0.02, -0.01, 1.0),
              t_sigma: float = 0.02, t_norm: float | None = 1.0, px_sigma: float = 0.5,
              outlier_frac: float = 0.30, t_dir_uniform: bool = False) -> Pair:
    """One KITTI-shaped (cfg1) frame pair; with EuRoC arguments, a cfg2 pair."""
    rng = np.random.default_rng(seed)
    W, H = size
    px0 = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], axis=1)
    d0 = rng.uniform(depth[0], depth[1], n)
    R = rodrigues(rng.normal(0.0, rot_sigma, 3))
    if t_dir_uniform:
        t = rng.normal(0.0, 1.0, 3)
        t /= np.linalg.norm(t)
    else:
        t = np.asarray(t_mean, dtype=np.float64) + rng.normal(0.0, t_sigma, 3)
    if t_norm is not None:
        t = t / np.linalg.norm(t) * t_norm
    X0 = _backproject(K, px0, d0)
    # the camera moves forward by t, so points move by -t in the camera frame:
    # X1 = R X0 + t_pt with t_pt the *point* translation recoverPose reports.
    t_pt = -t
    X1 = X0 @ R.T + t_pt
    px1 = _project(K, X1) + rng.normal(0.0, px_sigma, (n, 2))
    desc0 = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    desc1 = flip_bits(desc0, rng)
    out = rng.random(n) < outlier_frac
    out |= (X1[:, 2] <= 0.1)
    n_out = int(out.sum())
    px1[out] = np.stack([rng.uniform(0, W, n_out), rng.uniform(0, H, n_out)], axis=1)
    desc1[out] = rng.integers(0, 256, size=(n_out, 32), dtype=np.uint8)
    perm = rng.permutation(n)
    inv = np.empty(n, dtype=np.int64)
    inv[perm] = np.arange(n)
    gt = np.where(out, -1, inv).astype(np.int32)
    return Pair(kp0=px0.astype(np.float32), kp1=px1[perm].astype(np.float32),
                desc0=desc0, desc1=np.ascontiguousarray(desc1[perm]),
                K=K.copy(), R=R, t=t_pt / np.linalg.norm(t_pt), gt_match=gt)


def make_kitti_pair(index: int = 0, n: int = 2000) -> Pair:
    return make_pair(seed_for(1, index), n=n)


def make_euroc_pair(index: int = 0, n: int = 1500) -> Pair:
    return make_pair(seed_for(2, index), n=n, K=EUROC_K, size=EUROC_SIZE, depth=(1.0, 8.0),
                     rot_sigma=0.02, t_norm=0.05, px_sigma=0.3, outlier_frac=0.25,
                     t_dir_uniform=True)


@dataclasses.dataclass
class Sequence:
    """cfg3: F frames x n keypoints; pair i = (frame i, frame i+1)."""
    kps: np.ndarray        # (F, n, 2) f32
    descs: np.ndarray      # (F, n, 32) u8
    K: np.ndarray
    R: np.ndarray          # (F-1, 3, 3) ground-truth point rotations
    t: np.ndarray          # (F-1, 3)   ground-truth unit point translations

    @property
    def n_frames(self) -> int:
        return self.kps.shape[0]

    @property
    def n_pairs(self) -> int:
        return self.kps.shape[0] - 1


def make_sequence(n_frames: int = 4541, n: int = 2000, seed: int = seed_for(3, 0),
                  K: np.ndarray = KITTI_K, size=KITTI_SIZE, depth=(5.0, 60.0),
                  px_sigma: float = 0.5, outlier_frac: float = 0.30, step=(0.6, 1.2)) -> Sequence:
    """KITTI seq-00-length synthetic run with a smooth random-walk trajectory.

    Frame f+1 re-observes ~70 % of frame f's landmarks (re-projected under the frame
    motion, pixel noise, descriptor bit flips) and replaces the rest, plus whatever left
    the image, with fresh random features; the order is shuffled per frame.
    """
    rng = np.random.default_rng(seed)
    W, H = size
    kps = np.empty((n_frames, n, 2), dtype=np.float32)
    descs = np.empty((n_frames, n, 32), dtype=np.uint8)
    Rs = np.empty((max(n_frames - 1, 0), 3, 3))
    ts = np.empty((max(n_frames - 1, 0), 3))
    px = np.stack([rng.uniform(0, W, n), rng.uniform(0, H, n)], axis=1)
    dep = rng.uniform(depth[0], depth[1], n)
    desc = rng.integers(0, 256, size=(n, 32), dtype=np.uint8)
    kps[0] = px
    descs[0] = desc
    w = np.zeros(3)
    v = np.array([0.02, -0.01, 1.0])
    for f in range(1, n_frames):
        w = 0.9 * w + rng.normal(0.0, 0.004, 3)          # smooth yaw/pitch/roll walk
        v = 0.95 * v + 0.05 * np.array([0.02, -0.01, 1.0]) + rng.normal(0.0, 0.01, 3)
        R = rodrigues(w)
        t_pt = -v / np.linalg.norm(v) * rng.uniform(step[0], step[1])
        X1 = _backproject(K, px, dep) @ R.T + t_pt
        ok = X1[:, 2] > 0.5
        p1 = np.full((n, 2), -1.0)
        p1[ok] = _project(K, X1[ok])
        p1 += rng.normal(0.0, px_sigma, (n, 2))
        keep = ok & (rng.random(n) >= outlier_frac)
        keep &= (p1[:, 0] >= 0) & (p1[:, 0] < W) & (p1[:, 1] >= 0) & (p1[:, 1] < H)
        nd = flip_bits(desc, rng)
        n_new = int((~keep).sum())
        p1[~keep] = np.stack([rng.uniform(0, W, n_new), rng.uniform(0, H, n_new)], axis=1)
        nd[~keep] = rng.integers(0, 256, size=(n_new, 32), dtype=np.uint8)
        ndep = np.where(keep, X1[:, 2], rng.uniform(depth[0], depth[1], n))
        perm = rng.permutation(n)
        px, dep, desc = p1[perm], ndep[perm], np.ascontiguousarray(nd[perm])
        kps[f] = px
        descs[f] = desc
        Rs[f - 1] = R
        ts[f - 1] = t_pt / np.linalg.norm(t_pt)
    return Sequence(kps=kps, descs=descs, K=K.copy(), R=Rs, t=ts)


# --------------------------------------------------------------------------------------
# LM fixtures modelled on sequence.hpp:10-159 (seeded instead of srand(time(0)))
# --------------------------------------------------------------------------------------

def _rot_xyz(a, b, c):
    ca, sa, cb, sb, cc, sc = np.cos(a), np.sin(a), np.cos(b), np.sin(b), np.cos(c), np.sin(c)
    Rx = np.array([[1, 0, 0], [0, ca, -sa], [0, sa, ca]])
    Ry = np.array([[cb, 0, sb], [0, 1, 0], [-sb, 0, cb]])
    Rz = np.array([[cc, -sc, 0], [sc, cc, 0], [0, 0, 1]])
    return Rx @ Ry @ Rz


def gen_T(rng) -> np.ndarray:
    """sequence.hpp:10-29 -- rotation +-30 deg per axis, t in [-2,2]^3 with t_z >= 0."""
    ang = 2.0 * (0.5 - rng.random(3)) * np.pi / 6
    T = np.eye(4)
    T[:3, :3] = _rot_xyz(*ang)
    t = 2.0 * rng.uniform(-1, 1, 3)
    t[2] = abs(t[2])
    T[:3, 3] = t
    return T


def T_noise(rng, rot: float = 1e-1, tr: float = 1e-1) -> np.ndarray:
    """sequence.hpp:39-50 -- +-0.05 rad per axis, +-0.1 translation."""
    ang = rot * (0.5 - rng.random(3))
    T = np.eye(4)
    T[:3, :3] = _rot_xyz(*ang)
    T[:3, 3] = tr * rng.uniform(-1, 1, 3)
    return T


def gen_points(rng, N: int, T: np.ndarray):
    """sequence.hpp:64-104 -- view-0 z > 0, rejected until view-1 depth > 10."""
    R, t = T[:3, :3], T[:3, 3]
    mag = np.linalg.norm(t)
    X = np.zeros((N, 3))
    p = np.zeros((N, 3))
    p_ = np.zeros((N, 3))
    for i in range(N):
        while True:
            x = 10 * mag * rng.uniform(-1, 1, 3)
            if x[2] < 0:
                x[2] = -x[2] + 10.0
            x_ = R @ x + t
            if x_[2] > 10.0:
                break
        X[i] = x
        p[i] = x / x[2]
        p_[i] = x_ / x_[2]
    return X, p, p_


def compose_rep(Ts, z0: int, z1: int) -> np.ndarray:
    """sequence.hpp:140-149 -- product of the chain between zeta z0 and z1 (either order)."""
    T = np.eye(4)
    if z0 <= z1:
        for j in range(z0, z1 + 1):
            T = Ts[j] @ T
    else:
        for j in range(z0, z1 - 1, -1):
            T = np.linalg.inv(Ts[j]) @ T
    return T


def gen_scene_sequence(seed: int, N: int, n_zeta: int, reps, noise_rot=1e-1, noise_tr=1e-1):
    """sequence.hpp:106-159.  Returns Ts, T0s (n_zeta,4,4) and pr, p_r (n_rep,N,3)."""
    rng = np.random.default_rng(seed)
    Ts = np.stack([gen_T(rng) for _ in range(n_zeta)])
    T0s = np.stack([Ts[i] @ T_noise(rng, noise_rot, noise_tr) for i in range(n_zeta)])
    pr = np.zeros((len(reps), N, 3))
    p_r = np.zeros((len(reps), N, 3))
    for i, (z0, z1) in enumerate(reps):
        _, pr[i], p_r[i] = gen_points(rng, N, compose_rep(Ts, z0, z1))
    return Ts, T0s, pr, p_r


# ---- kitti_ba windows: synthetic reprojection map ---------------------------------------------
def make_reprojs(seed: int, num_frames: int, window, K: np.ndarray = KITTI_K, n_pts: int = 48, stereo: bool = False,
                 few_points_at=(), baseline: float = 0.54, pose_noise: float = 2e-3):
    """A `reprojs` map like the one robust_ass / robust_ass_stereo fill (kitti_ba.cpp:178-582) for a
    synthetic drive: smooth forward motion, landmarks 6..40 m ahead, pixel noise 0.2, and per
    consecutive node pair the relative pose with a unit translation (what recoverPose returns) plus
    a little noise.  Mono nodes are frames; stereo nodes are 2f (left) / 2f+1 (right) with the
    right camera `baseline` to the side and weight 0 on the left->right reprojection (frozen
    extrinsics, kitti_ba.cpp:171-172).  Keys in `few_points_at` get fewer than 32 points."""
    from .ba import Reproj, expand_stereo_window

    rng = np.random.default_rng(seed)
    nodes = 2 * num_frames if stereo else num_frames
    # world pose of every node (camera-from-world), frames move ~1 m forward with small rotations
    cam_T = [np.eye(4)]
    for _ in range(num_frames - 1):
        d = np.eye(4)
        d[:3, :3] = rodrigues(rng.normal(0, 0.01, 3))
        d[:3, 3] = np.array([0.02, -0.01, -1.0]) + rng.normal(0, 0.02, 3)      # point transform: scene moves backwards
        cam_T.append(d @ cam_T[-1])
    node_T = []
    for f in range(num_frames):
        node_T.append(cam_T[f])
        if stereo:
            e = np.eye(4)
            e[0, 3] = -baseline
            node_T.append(e @ cam_T[f])
    win = expand_stereo_window(window) if stereo else list(window)
    step = 2 if stereo else 1
    span = max(max(a, b) for a, b in win)
    keys = set()
    for i in range(0, num_frames):
        if step * i + span >= nodes:
            break
        for a, b in win:
            keys.add((step * i + a, step * i + b))
    lo = min(min(a, b) for a, b in win)
    for j in range(nodes - 1):
        keys.add((j, j + 1))
    out = {}
    for (i0, i1) in sorted(keys):
        rel = node_T[i1] @ np.linalg.inv(node_T[i0])                            # point transform i0 -> i1
        n = n_pts if (i0, i1) not in few_points_at else 20
        X0 = np.column_stack([rng.uniform(-8, 8, n), rng.uniform(-2, 2, n), rng.uniform(6, 40, n)])
        X1 = (rel[:3, :3] @ X0.T).T + rel[:3, 3]
        p0 = _project(K, X0) + rng.normal(0, 0.2, (n, 2))
        p1 = _project(K, X1) + rng.normal(0, 0.2, (n, 2))
        t = rel[:3, 3] / max(np.linalg.norm(rel[:3, 3]), 1e-12)
        R = rel[:3, :3] @ rodrigues(rng.normal(0, pose_noise, 3))
        t = t + rng.normal(0, pose_noise, 3)
        w = 0.0 if (stereo and i0 % 2 == 0 and i1 == i0 + 1) else 1.0
        out[(i0, i1)] = Reproj(p0.astype(np.float32), p1.astype(np.float32), R, t, w)
    return out


def corner_scene(rows: int, cols: int, seed: int) -> np.ndarray:
    """An 8-bit test image with corners at every scale for the detector / descriptor front end (FAST, ORB): smooth random
    background, filled rectangles of random grey, a little noise (no OpenCV needed)."""
    rng = np.random.default_rng(seed)
    coarse = rng.integers(60, 200, (rows // 8 + 3, cols // 8 + 3)).astype(np.float64)
    yy = np.arange(rows) / 8.0
    xx = np.arange(cols) / 8.0
    y0, x0 = yy.astype(int), xx.astype(int)
    fy, fx = (yy - y0)[:, None], (xx - x0)[None, :]
    img = (coarse[y0][:, x0] * (1 - fy) * (1 - fx) + coarse[y0 + 1][:, x0] * fy * (1 - fx) +
           coarse[y0][:, x0 + 1] * (1 - fy) * fx + coarse[y0 + 1][:, x0 + 1] * fy * fx)
    for _ in range(max(8, rows * cols // 600)):
        x, y = int(rng.integers(0, cols)), int(rng.integers(0, rows))
        w, h = (int(v) for v in rng.integers(4, 40, 2))
        img[y:y + h, x:x + w] = float(rng.integers(0, 256))
    img += rng.normal(0, 3, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)
