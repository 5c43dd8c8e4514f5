"""N4 front end throughput: FAST(40) + pyramidal LK over a KITTI-sized synthetic sequence through the host API
(images uploaded per call), next to the same cv2 calls on the host cores.   python tools/frontend_bench.py [frames]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rng = np.random.default_rng(1)
import cv2
base = cv2.normalize(cv2.GaussianBlur(rng.integers(0, 256, (460, 1400)).astype(np.uint8), (0, 0), 1.6), None, 0, 255, cv2.NORM_MINMAX)
frames = []
for k in range(n):
    s = 1.0 + 0.004 * (k % 16)
    M = np.array([[s, 0, -620 * (s - 1) + 0.5 * (k % 16)], [0, s, -190 * (s - 1)]], np.float32)
    frames.append(cv2.warpAffine(base, M, (1400, 460))[40:416, 60:1301].copy())
frames = np.stack(frames)
ctx = api.Context(0)
out = {"frames": n, "size": [376, 1241]}
for rep in range(3):
    t0 = time.perf_counter(); det = api.fastDetect(frames[:-1], 40, True, max_keypoints=8192, ctx=ctx); t1 = time.perf_counter()
    pts = [d[0] for d in det]
    nxt, st = api.trackSequenceLK(frames, pts, ctx=ctx); t2 = time.perf_counter()
out["gpu_fast_ms_per_frame"] = (t1 - t0) * 1e3 / (n - 1)
out["gpu_lk_ms_per_pair"] = (t2 - t1) * 1e3 / (n - 1)
out["mean_keypoints"] = float(np.mean([len(p) for p in pts]))
out["tracked_frac"] = float(np.mean([s.mean() for s in st]))
cv2.setNumThreads(0)
d = cv2.FastFeatureDetector_create(40)
m = min(n - 1, 16)
t0 = time.perf_counter()
cp = [np.array([k.pt for k in d.detect(frames[i], None)], np.float32).reshape(-1, 2) for i in range(m)]
t1 = time.perf_counter()
for i in range(m):
    cv2.calcOpticalFlowPyrLK(frames[i], frames[i + 1], cp[i], None)
t2 = time.perf_counter()
out["cv2_fast_ms_per_frame_1thread"] = (t1 - t0) * 1e3 / m
out["cv2_lk_ms_per_pair_1thread"] = (t2 - t1) * 1e3 / m
print(json.dumps(out))
