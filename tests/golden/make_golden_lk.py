"""Golden vectors for the LK tracker: cv2.calcOpticalFlowPyrLK (defaults, as at kitti_E.cpp:79-84) on synthetic image
pairs, run here with cv2 4.13.0; the outputs travel as tests/golden/lk.npz.   python tests/golden/make_golden_lk.py"""
import os

import cv2
import numpy as np


def pairs():
    rng = np.random.default_rng(77)
    out = {}
    base = cv2.GaussianBlur(rng.integers(0, 256, (260, 340)).astype(np.uint8), (0, 0), 1.8)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    # 1: affine motion of a few pixels, window-sized borders exercised (points down to 3 px from the border)
    M = np.array([[1.012, 0.004, 2.6], [-0.005, 1.009, -1.9]], np.float32)
    out["affine"] = (base[20:220, 20:300].copy(), cv2.warpAffine(base, M, (340, 260))[20:220, 20:300].copy())
    # 2: large motion (needs the pyramid) plus noise
    M = np.array([[1.0, 0.0, 11.3], [0.0, 1.0, -6.8]], np.float32)
    t = cv2.warpAffine(base, M, (340, 260))[30:200, 40:280].astype(np.int32) + rng.integers(-6, 7, (170, 240))
    out["shift_noise"] = (base[30:200, 40:280].copy(), np.clip(t, 0, 255).astype(np.uint8))
    # 3: small image, only three pyramid levels exist (47 x 60 -> 24 x 30 -> stop); flat regions lose the track
    s = base[100:147, 100:160].copy()
    s[:, 40:] = 128
    out["small_flat"] = (s, np.roll(s, 1, axis=1))
    return out


def main():
    res = {}
    for name, (a, b) in pairs().items():
        kps = cv2.FastFeatureDetector_create(10).detect(a, None)
        pts = np.array([k.pt for k in kps], dtype=np.float32).reshape(-1, 2)[:600]
        rng = np.random.default_rng(len(pts))
        extra = np.stack([rng.uniform(0, a.shape[1] - 1, 40), rng.uniform(0, a.shape[0] - 1, 40)], axis=1).astype(np.float32)
        pts = np.concatenate([pts, extra])                    # sub-pixel points anywhere, borders included
        nxt, st, err = cv2.calcOpticalFlowPyrLK(a, b, pts, None)
        res["prev_" + name], res["next_" + name], res["pts_" + name] = a, b, pts
        res["out_" + name], res["status_" + name] = nxt.reshape(-1, 2), st.ravel()
        res["err_" + name] = np.where(st.ravel() == 1, err.ravel(), 0).astype(np.float32)     # cv2 leaves err of lost tracks undefined
        print(name, a.shape, len(pts), "tracked", int(st.sum()))
    res["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lk.npz"), **res)


if __name__ == "__main__":
    main()
