"""Census of the ORB restatement (oracle/orb.py) against live cv2 over random configurations: image size, nfeatures,
scaleFactor, nlevels, edgeThreshold, fastThreshold.  One JSON line: how many configurations (and keypoints / descriptors)
were identical.   python tests/orb_census.py [n_configs] > profiles/r2_orb_census.json      (CPU only; test infrastructure:
it imports the oracle)"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cv2
import numpy as np
from epivo_b200.orb_pattern import BIT_PATTERN_31
from oracle import orb as OO
from orb_util import cv2_orb, kps_array, scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(2026)
same = kp_total = refused = 0
bad = []
for it in range(n):
    rows, cols = int(rng.integers(40, 420)), int(rng.integers(40, 700))
    nf = int(rng.choice([30, 200, 1000, 3000, 10000]))
    sc = float(np.float32(rng.choice([1.1, 1.2, 1.25, 1.3, 1.5, 1.7, 2.0])))
    nl = int(rng.integers(1, 13))
    edge = int(rng.integers(15, 40))
    thr = int(rng.integers(5, 60))
    img = scene(rows, cols, 1000 + it)
    if it % 5 == 0:
        img = cv2.GaussianBlur(img, (0, 0), 1.2)                      # softer corners, more score ties
    sizes = OO.layer_sizes(rows, cols, OO.layer_scales(nl, sc))
    try:
        ref_k, ref_d = cv2_orb(cv2, img, nf, sc, nl, edge, thr)
    except cv2.error:
        # a pyramid level of size 0: OpenCV asserts in resize; epivo_orb_detect_and_compute refuses the same configurations
        assert min(min(r, c) for r, c in sizes) < 1, (rows, cols, sc, nl)
        refused += 1
        continue
    assert min(min(r, c) for r, c in sizes) >= 1
    k, d = OO.detect_and_compute(img, BIT_PATTERN_31, nf, sc, nl, edge, thr)
    ok = k.shape[0] == ref_k.shape[0] and np.array_equal(kps_array(k), ref_k) and np.array_equal(d, ref_d)
    same += int(ok)
    kp_total += len(ref_k)
    if not ok:
        bad.append({"rows": rows, "cols": cols, "nfeatures": nf, "scale": sc, "nlevels": nl, "edge": edge, "fast": thr,
                    "cv2_keypoints": int(len(ref_k)), "oracle_keypoints": int(k.shape[0])})
print(json.dumps({"configs": n, "refused_by_cv2_and_by_us_empty_level": refused, "identical": same, "keypoints_compared": kp_total, "cv2": cv2.__version__, "differing": bad}))
