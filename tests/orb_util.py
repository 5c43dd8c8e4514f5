"""Shared by the ORB tests: a synthetic scene with corners at every scale, and cv2's result as plain arrays."""
import numpy as np

FIELDS = ("x", "y", "size", "angle", "response", "octave")


def scene(rows: int, cols: int, seed: int) -> np.ndarray:
    """Smooth random background, filled rectangles of random grey, a little noise (no OpenCV needed)."""
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


def cv2_orb(cv2, img, nfeatures=10000, scale=1.2, nlevels=8, edge=15, fast_thr=20):
    """(kps (n, 6) float32 in FIELDS order, desc (n, 32) uint8) of the reference's ORB configuration (kitti_ba.cpp:128)."""
    orb = cv2.ORB_create(nfeatures, scale, nlevels, edge, 0, 2, cv2.ORB_FAST_SCORE, 31, fast_thr)
    kp = orb.detect(img, None)
    kp, desc = orb.compute(img, kp)
    ref = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave) for k in kp], np.float32).reshape(-1, 6)
    return ref, (np.zeros((0, 32), np.uint8) if desc is None else desc)


def kps_array(kps) -> np.ndarray:
    """epivo keypoints (structured array) or the oracle's (n, 7) rows -> (n, 6) float32 in FIELDS order."""
    if getattr(kps, "dtype", None) is not None and kps.dtype.names:
        return np.stack([kps[f].astype(np.float32) for f in FIELDS], axis=1).reshape(-1, 6)
    return np.asarray(kps, np.float32).reshape(-1, 7)[:, :6]
