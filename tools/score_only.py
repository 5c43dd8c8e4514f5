"""K3 alone on the largest cfg4 point (M = 65536 models x N = 20000 correspondences), for an ncu capture of
score_count_kernel: `ncu --set full -k regex:score_count -s 2 -c 1 python tools/score_only.py`."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api, synth

ctx = api.Context(0)
N, M = int(os.environ.get("N", "20000")), int(os.environ.get("M", "65536"))
p = synth.make_pair(synth.seed_for(4, N), n=N, outlier_frac=0.5)
Kf = p.K.astype(np.float32)
rng = np.random.default_rng(1)
c0 = p.kp0
c1 = np.ascontiguousarray(p.kp1[np.where(p.gt_match >= 0, p.gt_match, rng.integers(0, N, N))])
# models of realistic scale: the true E perturbed
E0 = p.E / np.linalg.norm(p.E) if hasattr(p, "E") else None
if E0 is None:
    tx = np.array([[0, -p.t[2], p.t[1]], [p.t[2], 0, -p.t[0]], [-p.t[1], p.t[0], 0]])
    E0 = tx @ p.R
    E0 /= np.linalg.norm(E0)
models = E0.reshape(1, 9) + 0.002 * rng.standard_normal((M, 9))
models /= np.linalg.norm(models, axis=1, keepdims=True)
for _ in range(4):
    t0 = time.perf_counter()
    cnt = api.scoreSampson(models, c0, c1, Kf, 1.0, ctx=ctx, medians=False)[0]
    dt = time.perf_counter() - t0
print("host-API rate %.1f G hypothesis-points/s, best count %d" % (M * N / dt / 1e9, cnt.max()))
