"""Shared by the ORB tests: a synthetic scene with corners at every scale, and cv2's result as plain arrays."""
import numpy as np

from epivo_b200 import synth

FIELDS = ("x", "y", "size", "angle", "response", "octave")


scene = synth.corner_scene


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
