"""N4 (host side): the ground-truth feeds of the reference drivers -- KITTI pose files, EuRoC ground-truth
CSVs with quaternion poses and timestamp association -- and the per-pair ground-truth step lengths the
drivers use to scale the unit-norm translation of every pair (kitti_E.cpp:216-222, euroc_E.cpp:300-304).
These produce the `scales` argument of `SequencePipeline.cloud` / `shard.chain_poses`.

Image loading, FAST/ORB detection and LK tracking stay out of scope (the benchmark inputs are keypoint /
descriptor sets); everything here is plain numpy on the host.
"""
from __future__ import annotations

import numpy as np


def load_kitti_poses(path: str) -> np.ndarray:
    """KITTI odometry `poses/NN.txt`: 12 space-separated numbers per line = a 3x4 row-major [R|t]
    (load_csv, kitti_E.cpp:18-34; reshaped to 4x3 and transposed at kitti_E.cpp:207-214) -> (n, 4, 4)."""
    rows = np.loadtxt(path, dtype=np.float64, ndmin=2)
    if rows.shape[1] != 12:
        raise ValueError(f"{path}: expected 12 numbers per line, got {rows.shape[1]}")
    T = np.tile(np.eye(4), (rows.shape[0], 1, 1))
    T[:, :3, :4] = rows.reshape(-1, 3, 4)
    return T


def gt_point_transforms(poses: np.ndarray) -> np.ndarray:
    """dT_i = (pT_i^-1 T_{i+1})^-1: the ground-truth point transform of pair i (kitti_E.cpp:216-217)."""
    poses = np.asarray(poses, dtype=np.float64)
    return np.stack([np.linalg.inv(np.linalg.inv(poses[i]) @ poses[i + 1]) for i in range(len(poses) - 1)])


def gt_scales(poses: np.ndarray) -> np.ndarray:
    """`scale = dT.block<3,1>(0,3).norm()` per pair (kitti_E.cpp:219)."""
    return np.linalg.norm(gt_point_transforms(poses)[:, :3, 3], axis=1)


def quat_to_R(q) -> np.ndarray:
    """(qw, qx, qy, qz) -> rotation matrix, normalising first (euroc_E.cpp:66-86)."""
    qw, qx, qy, qz = (float(v) for v in q)
    n = 1.0 / np.sqrt(qx * qx + qy * qy + qz * qz + qw * qw)
    qw, qx, qy, qz = qw * n, qx * n, qy * n, qz * n
    return np.array([[1.0 - 2.0 * qy * qy - 2.0 * qz * qz, 2.0 * qx * qy - 2.0 * qz * qw, 2.0 * qx * qz + 2.0 * qy * qw],
                     [2.0 * qx * qy + 2.0 * qz * qw, 1.0 - 2.0 * qx * qx - 2.0 * qz * qz, 2.0 * qy * qz - 2.0 * qx * qw],
                     [2.0 * qx * qz - 2.0 * qy * qw, 2.0 * qy * qz + 2.0 * qx * qw, 1.0 - 2.0 * qx * qx - 2.0 * qy * qy]])


def load_euroc_groundtruth(path: str) -> np.ndarray:
    """EuRoC `state_groundtruth_estimate0/data.csv`: header line, then comma-separated rows
    timestamp, px, py, pz, qw, qx, qy, qz, ... (load_csv with the first line skipped, euroc_E.cpp:23-44)."""
    return np.loadtxt(path, delimiter=",", skiprows=1, dtype=np.float64, ndmin=2)


def load_euroc_image_timestamps(path: str) -> np.ndarray:
    """EuRoC `cam0/data.csv`: header line, then `timestamp,filename`; the reference keeps the first column
    (load_fns, euroc_E.cpp:47-63) and later parses it with stod."""
    out = []
    with open(path) as f:
        next(f, None)
        for line in f:
            cell = line.split(",")[0].strip()
            if cell:
                out.append(float(cell))
    return np.array(out, dtype=np.float64)


EUROC_TS_TOLERANCE = 5760512 - 760576      # ns, euroc_E.cpp:228: just under the 5 ms period of the 200 Hz ground truth


def associate_euroc(gt: np.ndarray, image_ts: np.ndarray, tol: float = EUROC_TS_TOLERANCE) -> np.ndarray:
    """For every image timestamp the first ground-truth row within `tol` of it (euroc_E.cpp:226-246, which
    scans forward from a guessed row and takes the first hit) -> (n, 4, 4) body poses [R(q) | p]; raises if
    an image has no match (the reference asserts cnt_found == 2)."""
    ts = gt[:, 0]
    T = np.tile(np.eye(4), (len(image_ts), 1, 1))
    for i, t in enumerate(image_ts):
        j = int(np.searchsorted(ts, t - tol, side="right"))        # first row with ts > t - tol
        if j >= len(ts) or not abs(ts[j] - t) < tol:
            raise ValueError(f"no ground-truth row within {tol} ns of image timestamp {t:.0f}")
        T[i, :3, :3] = quat_to_R(gt[j, 4:8])
        T[i, :3, 3] = gt[j, 1:4]
    return T


def euroc_gt_scales(body_poses: np.ndarray, T_DC: np.ndarray) -> np.ndarray:
    """`dT = ((pT T_DC)^-1 (T T_DC))^-1`, scale = |dT.t| (euroc_E.cpp:300-301): camera-frame step lengths."""
    cam = np.asarray(body_poses) @ np.asarray(T_DC, dtype=np.float64)
    return gt_scales(cam)


# EuRoC cam0 calibration as euroc_E.cpp:88-104 hard-codes it (sensor.yaml intrinsics + the stereo rectification it uses)
EUROC_CAM0_K = np.array([[458.654, 0.0, 367.215], [0.0, 457.296, 248.375], [0.0, 0.0, 1.0]])
EUROC_CAM0_DIST = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
EUROC_CAM0_RECT = np.array([[0.999966347530033, -0.001422739138722922, 0.008079580483432283],
                            [0.001365741834644127, 0.9999741760894847, 0.007055629199258132],
                            [-0.008089410156878961, -0.007044357138835809, 0.9999424675829176]])
EUROC_CAM0_PROJ = np.array([[435.2046959714599, 0.0, 367.4517211914062, 0.0],
                            [0.0, 435.2046959714599, 252.2008514404297, 0.0],
                            [0.0, 0.0, 1.0, 0.0]])


def undistort_rectify_maps(K, dist, R, P, size):
    """The fixed-point maps `initUndistortRectifyMap(cam, dist, rect, proj, Size(w, h), map1.type(), map1, map2)` makes
    once before the frame loop (euroc_E.cpp:105-113; m1type 0 selects CV_16SC2 + CV_16UC1): (map_xy (h, w, 2) int16,
    map_frac (h, w) uint16), the input of `api.remap` / `epivo_remap`.  Radial-tangential model k1 k2 p1 p2 [k3]."""
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    d = np.zeros(5)
    dv = np.ravel(np.asarray(dist, dtype=np.float64))
    if len(dv) > 5:
        raise ValueError("only k1 k2 p1 p2 [k3] are supported")
    d[:len(dv)] = dv
    k1, k2, p1, p2, k3 = d
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
    P = np.asarray(P, dtype=np.float64)
    iR = np.linalg.inv(P[:3, :3] @ R)                       # destination pixel -> rectified ray
    w, h = size
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    W = iR[2, 0] * u + iR[2, 1] * v + iR[2, 2]
    x = (iR[0, 0] * u + iR[0, 1] * v + iR[0, 2]) / W
    y = (iR[1, 0] * u + iR[1, 1] * v + iR[1, 2]) / W
    x2, y2, xy2 = x * x, y * y, 2 * x * y
    r2 = x2 + y2
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    mu = K[0, 0] * (x * kr + p1 * xy2 + p2 * (r2 + 2 * x2)) + K[0, 2]          # distorted source position
    mv = K[1, 1] * (y * kr + p1 * (r2 + 2 * y2) + p2 * xy2) + K[1, 2]
    iu, iv = np.rint(mu * 32).astype(np.int64), np.rint(mv * 32).astype(np.int64)   # 1/32-pixel fixed point
    xy = np.stack([np.clip(iu >> 5, -32768, 32767), np.clip(iv >> 5, -32768, 32767)], axis=2).astype(np.int16)
    return xy, ((iv & 31) * 32 + (iu & 31)).astype(np.uint16)
