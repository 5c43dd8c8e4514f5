"""N4 front end, ORB throughput: epivo_orb_detect_and_compute (kitti_ba.cpp:128's ORB::create(10000, 1.2f, 8, 15, 0, 2,
FAST_SCORE)) over KITTI-sized synthetic frames through the host API (frames uploaded, keypoints and descriptors
downloaded inside the timed region), next to cv2's ORB on the host cores.   python tools/orb_bench.py [frames]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from epivo_b200 import api
from orb_util import scene

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
frames = np.stack([scene(376, 1241, 500 + i) for i in range(n)])
ctx = api.Context(0)
out = {"frames": n, "size": [376, 1241], "config": "ORB(10000, 1.2, 8, 15, 0, 2, FAST_SCORE)"}
for batch in (1, 8, n):
    best = 1e9
    for rep in range(4):
        t0 = time.perf_counter()
        for i in range(0, n, batch):
            res = api.orbDetectAndCompute(frames[i:i + batch], 10000, max_keypoints=12000, ctx=ctx)
        best = min(best, time.perf_counter() - t0)
    out["gpu_ms_per_frame_batch%d" % batch] = best * 1e3 / n
res = api.orbDetectAndCompute(frames, 10000, max_keypoints=12000, ctx=ctx)
out["mean_keypoints"] = float(np.mean([len(r[0]) for r in res]))
out["kernel_launches_per_call"] = int(ctx.launch_count)
c0 = ctx.launch_count
api.orbDetectAndCompute(frames, 10000, max_keypoints=12000, ctx=ctx)
out["kernel_launches_per_call"] = int(ctx.launch_count - c0)
if "--no-cv2" not in sys.argv:
    import cv2
    m = min(n, 8)
    for threads in (1, 0):
        cv2.setNumThreads(threads)
        orb = cv2.ORB_create(10000, 1.2, 8, 15, 0, 2, cv2.ORB_FAST_SCORE)
        t0 = time.perf_counter()
        for i in range(m):
            kp = orb.detect(frames[i], None)
            kp, d = orb.compute(frames[i], kp)
        out["cv2_ms_per_frame_%s" % ("1thread" if threads == 1 else "allthreads")] = (time.perf_counter() - t0) * 1e3 / m
    kp, d = orb.compute(frames[0], orb.detect(frames[0], None))
    out["identical_to_cv2_frame0"] = bool(np.array_equal(d, res[0][1]) and np.array_equal(np.array([k.pt for k in kp], np.float32), np.stack([res[0][0]["x"], res[0][0]["y"]], 1)))
print(json.dumps(out))
