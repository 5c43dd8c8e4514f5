"""Writers / readers of the text files the reference drivers emit and its Pangolin viewers replay
(kitti_E.cpp:257-286, euroc_E.cpp:351-372; read back by cloud_pango.py:25-39 with
`np.fromfile(path, sep=' ')`, which accepts any whitespace):

    pts.cld                       one "x y z" line per cloud point, each followed by a blank line
    lims                          cloud-point counts before each pair, space separated, one line
    kitti.T / kitti.GT / euroc.T  4x4 matrices, four rows each, blocks separated by a blank line
"""
from __future__ import annotations

import numpy as np


def write_cloud(path: str, points: np.ndarray) -> None:
    """pts.cld (kitti_E.cpp:258-264: `pt_cloud << X[i].transpose() << "\\n\\n"`)."""
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    with open(path, "w") as f:
        for p in pts:
            f.write("%.17g %.17g %.17g\n\n" % (p[0], p[1], p[2]))


def write_limits(path: str, limits) -> None:
    """lims (kitti_E.cpp:266-271: `lims << limits[i] << " "`)."""
    with open(path, "w") as f:
        for v in np.asarray(limits).ravel():
            f.write("%d " % int(v))


def write_poses(path: str, poses: np.ndarray) -> None:
    """kitti.T / kitti.GT / euroc.T (kitti_E.cpp:273-286): 4x4 blocks separated by a blank line."""
    with open(path, "w") as f:
        for T in np.asarray(poses, dtype=np.float64).reshape(-1, 4, 4):
            for r in range(4):
                f.write(" ".join("%.17g" % float(v) for v in T[r]) + "\n")
            f.write("\n")


def read_cloud(path: str) -> np.ndarray:
    """As cloud_pango.py:25-27 reads it."""
    return np.fromfile(path, sep=" ").reshape(-1, 3)


def read_limits(path: str) -> np.ndarray:
    return np.fromfile(path, sep=" ").astype(np.int64)


def read_poses(path: str) -> np.ndarray:
    """As cloud_pango.py:32-34 reads it."""
    return np.fromfile(path, sep=" ").reshape(-1, 4, 4)
