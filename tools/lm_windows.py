"""GPU tuning helper: time the windowed LM kernel on the cfg5 shape (n_zeta = 10, 20 reps x 250) and the shipped
kitti_ba stereo shape (n_zeta = 4, 9 reps x 32)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api, synth

ctx = api.Context(0)
REPS10 = [(i, i) for i in range(10)] + [(0, i) for i in range(10)]
STEREO = [(0, 1), (1, 1), (0, 0), (0, 3), (1, 3), (0, 0), (2, 3), (3, 3), (2, 2)]
B5 = int(os.environ.get("B5", "504"))
for name, nz, reps, N, B in [("cfg5", 10, REPS10, 250, B5), ("stereo_ws3", 4, STEREO, 32, max(1, B5 * 2270 // 504))]:
    data = [synth.gen_scene_sequence(500 + b, N, nz, reps) for b in range(32)]
    T0 = np.stack([data[b % 32][1] for b in range(B)])
    pr = np.stack([data[b % 32][2] for b in range(B)])
    p_r = np.stack([data[b % 32][3] for b in range(B)])
    for delta in (1.0, 1e-5):
        f = lambda: api.Levenberg_Marquardt_batch(nz, 1e-8, reps, [1.0] * len(reps), 1e-2, T0, pr, p_r, huber_delta=delta, ctx=ctx)
        f()
        t0 = time.perf_counter()
        T, res, its = f()
        dt = time.perf_counter() - t0
        print(f"{name} delta={delta}: {B / dt:.0f} windows/s  ({dt * 1e3:.1f} ms, mean iters {its.mean():.1f})")
