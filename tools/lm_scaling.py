"""GPU tuning helper: how the windowed LM kernel's time splits between per-point work and per-iteration overhead."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api, synth
ctx = api.Context(0)
REPS10 = [(i, i) for i in range(10)] + [(0, i) for i in range(10)]
B = 296
for nz, reps, N in [(10, REPS10, 250), (10, REPS10, 64), (10, REPS10, 16), (10, [(i, i) for i in range(10)], 250),
                    (10, [(0, 9)] * 2, 250), (4, [(i, i) for i in range(4)] + [(0, i) for i in range(4)], 250)]:
    data = [synth.gen_scene_sequence(500 + b, N, nz, reps) for b in range(8)]
    T0 = np.stack([data[b % 8][1] for b in range(B)])
    pr = np.stack([data[b % 8][2] for b in range(B)])
    p_r = np.stack([data[b % 8][3] for b in range(B)])
    api.Levenberg_Marquardt_batch(nz, 1e-8, reps, [1.0] * len(reps), 1e-2, T0, pr, p_r, huber_delta=1.0, ctx=ctx)
