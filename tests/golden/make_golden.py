"""Generate the golden vectors under tests/golden/ from cv2 (OpenCV) in the build container.

    python tests/golden/make_golden.py

OpenCV is the un-vendored library the reference calls for matching, essential-matrix
estimation and pose recovery (kitti_ba.cpp:602,641,702,715; kitti_E.cpp:98,120;
kitti.cpp:98; euroc_E.cpp:202,251).  The outputs recorded here are those of
`cv2.__version__` (written into every file) on seeded synthetic inputs from
`epivo_b200.synth`; the oracle (oracle/oracle.py) and the CUDA path are both checked
against them.  The GPU box never runs this script.
"""
from __future__ import annotations

import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from epivo_b200 import synth  # noqa: E402


def _matches(ms):
    return np.array([(m.queryIdx, m.trainIdx, int(m.distance)) for m in ms], dtype=np.int32).reshape(-1, 3)


def golden_match():
    out = {"cv2_version": cv2.__version__}
    rng = np.random.default_rng(101)
    cases = {
        "rand": (rng.integers(0, 256, (257, 32), dtype=np.uint8), rng.integers(0, 256, (301, 32), dtype=np.uint8)),
        "ties": (rng.integers(0, 256, (200, 32), dtype=np.uint8) & 0x03,
                 rng.integers(0, 256, (190, 32), dtype=np.uint8) & 0x03),
        "one": (rng.integers(0, 256, (1, 32), dtype=np.uint8), rng.integers(0, 256, (7, 32), dtype=np.uint8)),
    }
    pr = synth.make_kitti_pair(0, n=600)
    cases["kitti600"] = (pr.desc0, pr.desc1)
    for name, (q, t) in cases.items():
        out[f"{name}_q"], out[f"{name}_t"] = q, t
        for norm in (cv2.NORM_HAMMING, cv2.NORM_HAMMING2):
            for cc in (False, True):
                out[f"{name}_n{norm}_cc{int(cc)}"] = _matches(cv2.BFMatcher(norm, cc).match(q, t))
            if t.shape[0] >= 2:
                knn = cv2.BFMatcher(norm).knnMatch(q, t, k=2)
                out[f"{name}_n{norm}_knn_idx"] = np.array([[m.trainIdx for m in r] for r in knn], dtype=np.int32)
                out[f"{name}_n{norm}_knn_dist"] = np.array([[int(m.distance) for m in r] for r in knn], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "match.npz"), **out)


def _matched_points(pr):
    ms = _matches(cv2.BFMatcher(cv2.NORM_HAMMING2, True).match(pr.desc0, pr.desc1))
    return pr.kp0[ms[:, 0]], pr.kp1[ms[:, 1]]


def golden_essential():
    out = {"cv2_version": cv2.__version__}
    cases = [("kitti", synth.make_kitti_pair(1, n=700)), ("kitti_b", synth.make_kitti_pair(2, n=701)),
             ("euroc", synth.make_euroc_pair(0, n=600))]
    calls = [("ransac10", cv2.RANSAC, 0.99, 1.0), ("ransac03", cv2.RANSAC, 0.99, 0.3),
             ("ransac005", cv2.RANSAC, 0.99, 0.05), ("lmeds", cv2.LMEDS, 0.99, 0.01)]
    for name, pr in cases:
        p0, p1 = _matched_points(pr)
        Kf = pr.K.astype(np.float32)
        out[f"{name}_p0"], out[f"{name}_p1"], out[f"{name}_K"] = p0, p1, Kf
        out[f"{name}_Rgt"], out[f"{name}_tgt"] = pr.R, pr.t
        for cname, method, prob, thr in calls:
            E, mask = cv2.findEssentialMat(p0, p1, Kf, method, prob, thr)
            out[f"{name}_{cname}_E"] = E
            out[f"{name}_{cname}_mask"] = mask.ravel()
            m = mask.ravel() == 1
            n, R, t, rm = cv2.recoverPose(E, p0[m], p1[m], Kf)
            out[f"{name}_{cname}_pose_n"] = np.int32(n)
            out[f"{name}_{cname}_R"], out[f"{name}_{cname}_t"] = R, t.ravel()
            out[f"{name}_{cname}_pose_mask"] = rm.ravel()
    # minimal (N == 5) calls return every solution of the 5-point solver
    pr = synth.make_kitti_pair(3, n=300)
    p0, p1 = _matched_points(pr)
    Kf = pr.K.astype(np.float32)
    for trial in range(6):
        sel = np.random.default_rng(500 + trial).choice(len(p0), 5, replace=False)
        E, _ = cv2.findEssentialMat(p0[sel], p1[sel], Kf, cv2.RANSAC, 0.99, 1.0)
        out[f"min{trial}_p0"], out[f"min{trial}_p1"], out[f"min{trial}_K"] = p0[sel], p1[sel], Kf
        out[f"min{trial}_E"] = np.zeros((0, 3)) if E is None else E
    # small even-N LMedS calls pin the median rule (upper-middle order statistic).  N >= 12 so that
    # the median is not one of the five (numerically zero) sample residuals plus noise.
    qi = np.arange(len(p0))
    for trial in range(8):
        rng = np.random.default_rng(700 + trial)
        n = int(rng.choice([12, 14, 16, 20, 30]))
        sel = rng.choice(len(qi), n, replace=False)
        E, mask = cv2.findEssentialMat(p0[sel], p1[sel], Kf, cv2.LMEDS, 0.99, 0.01)
        if E is None or E.shape != (3, 3):
            continue
        out[f"small{trial}_p0"], out[f"small{trial}_p1"], out[f"small{trial}_K"] = p0[sel], p1[sel], Kf
        out[f"small{trial}_E"], out[f"small{trial}_mask"] = E, mask.ravel()
    np.savez_compressed(os.path.join(HERE, "essential.npz"), **out)


def golden_callsites():
    """The remaining findEssentialMat call shapes of the reference, which differ from the ones above in `prob`
    (it drives RANSACUpdateNumIters): kitti_ba.cpp:232 RANSAC(0.95, 0.05), kitti_ba.cpp:1279-1285
    RANSAC(0.999, 0.3), kitti_ba.cpp:702 LMEDS(0.99, 0.1), euroc_E.cpp:205-207 RANSAC(0.99, 0.3) on the EuRoC
    camera.  Separate file, so that essential.npz stays byte-identical."""
    out = {"cv2_version": cv2.__version__}
    cases = [("kitti", synth.make_kitti_pair(11, n=500)), ("euroc", synth.make_euroc_pair(5, n=450))]
    calls = [("ransac095_005", cv2.RANSAC, 0.95, 0.05), ("ransac0999_03", cv2.RANSAC, 0.999, 0.3),
             ("lmeds_01", cv2.LMEDS, 0.99, 0.1), ("ransac099_03", cv2.RANSAC, 0.99, 0.3)]
    for name, pr in cases:
        p0, p1 = _matched_points(pr)
        Kf = pr.K.astype(np.float32)
        out[f"{name}_p0"], out[f"{name}_p1"], out[f"{name}_K"] = p0, p1, Kf
        for cname, method, prob, thr in calls:
            E, mask = cv2.findEssentialMat(p0, p1, Kf, method, prob, thr)
            out[f"{name}_{cname}_E"] = E
            out[f"{name}_{cname}_mask"] = mask.ravel()
            out[f"{name}_{cname}_args"] = np.array([method, prob, thr], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "essential_callsites.npz"), **out)


if __name__ == "__main__":
    if "--callsites-only" in sys.argv:
        golden_callsites()
        sys.exit(0)
    golden_match()
    golden_essential()
    golden_callsites()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))
