"""CPU restatement of the detector the reference's front end calls (TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU legs -- never by the product path).

cv::FastFeatureDetector (FAST-9/16) is an OpenCV dependency of the reference, not reference source:
    Ptr<FastFeatureDetector> detector = FastFeatureDetector::create(40);  detector->detect(src, kp0, Mat());
        kitti_E.cpp:71-74, kitti_ba.cpp:49,62 (threshold 40), kitti_ba.cpp:98,117-118 (default threshold 10).
OpenCV is un-vendored and un-versioned in the reference (compile_cv: `pkg-config opencv`); the algorithm restated here
is the published one of modules/features2d/src/fast.cpp (FAST_t<16>) and fast_score.cpp (cornerScore<16>), and it is
pinned against `cv2 4.13.0` -- live in tests/test_oracle_fast.py and through tests/golden/fast.npz.
"""
from __future__ import annotations

import numpy as np

# (dx, dy) of the 16-pixel circle in OpenCV's order (fast_score.cpp makeOffsets, patternSize 16)
CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
          (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def _arc9(mask: np.ndarray) -> np.ndarray:
    """mask: (16, h, w) bool -> (h, w) bool, true where 9 circularly contiguous entries are set."""
    out = np.zeros(mask.shape[1:], dtype=bool)
    for s in range(16):
        run = np.ones(mask.shape[1:], dtype=bool)
        for k in range(9):
            run &= mask[(s + k) & 15]
        out |= run
    return out


def fast_scores(img: np.ndarray, threshold: int):
    """(corner (rows, cols) bool, score (rows, cols) int) of FAST_t<16>: fast.cpp -- a pixel with a full circle
    (3 <= x < cols-3, 3 <= y < rows-3) is a corner when 9 contiguous circle pixels are all < v - t or all > v + t
    (the pairwise rejection cascade in front of that test never rejects one); score = cornerScore<16>."""
    img = np.asarray(img, dtype=np.uint8)
    rows, cols = img.shape
    corner = np.zeros((rows, cols), dtype=bool)
    score = np.zeros((rows, cols), dtype=np.int32)
    threshold = int(threshold)
    if not 0 <= threshold <= 255:      # beyond 8 bits cv2's vector path truncates the threshold and its scalar tail does not
        raise ValueError("threshold outside [0, 255]")
    if rows < 7 or cols < 7:
        return corner, score
    v = img[3:rows - 3, 3:cols - 3].astype(np.int32)
    d = np.stack([v - img[3 + dy:rows - 3 + dy, 3 + dx:cols - 3 + dx].astype(np.int32) for dx, dy in CIRCLE])
    c = _arc9(d > threshold) | _arc9(-d > threshold)
    # cornerScore<16>: a0 = max(t, max over the 16 arcs of 9 of min d); b0 = min(-a0, min over arcs of max d); -b0-1
    a0 = np.full(v.shape, threshold, dtype=np.int32)
    arcs_min, arcs_max = [], []
    for s in range(16):
        idx = [(s + k) & 15 for k in range(9)]
        arcs_min.append(d[idx].min(axis=0))
        arcs_max.append(d[idx].max(axis=0))
    a0 = np.maximum(a0, np.max(arcs_min, axis=0))
    b0 = np.minimum(-a0, np.min(arcs_max, axis=0))
    sc = -b0 - 1
    corner[3:rows - 3, 3:cols - 3] = c
    score[3:rows - 3, 3:cols - 3] = np.where(c, sc, 0)
    return corner, score


def fast_detect(img: np.ndarray, threshold: int = 10, nonmax: bool = True):
    """cv2.FastFeatureDetector_create(threshold, nonmax).detect(img): (pts (k, 2) float32 [x, y], response (k,) float32)
    in OpenCV's order (row by row, left to right).  With suppression a corner survives when its score is strictly
    greater than the scores of its 8 neighbours (0 for non-corners); without it the response is 0 (fast.cpp keeps the
    zero-initialised score row)."""
    corner, score = fast_scores(img, threshold)
    if nonmax:
        s = np.pad(score, 1)
        keep = corner.copy()
        rows, cols = score.shape
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx or dy:
                    keep &= score > s[1 + dy:1 + dy + rows, 1 + dx:1 + dx + cols]
    else:
        keep = corner
    ys, xs = np.nonzero(keep)                                   # row-major: OpenCV's order
    pts = np.stack([xs, ys], axis=1).astype(np.float32).reshape(-1, 2)
    resp = (score[ys, xs] if nonmax else np.zeros(len(xs))).astype(np.float32)
    return pts, resp
