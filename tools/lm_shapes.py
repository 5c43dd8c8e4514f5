"""GPU tuning helper: kernel time of the cfg5 windowed LM for a batch size and CTA / cluster shape.
   EPIVO_LM_SHAPE=192x64x4 python tools/lm_shapes.py 63"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api, synth
ctx = api.Context(0)
REPS10 = [(i, i) for i in range(10)] + [(0, i) for i in range(10)]
B = int(sys.argv[1]) if len(sys.argv) > 1 else 63
data = [synth.gen_scene_sequence(500 + b, 250, 10, REPS10) for b in range(16)]
T0 = np.stack([data[b % 16][1] for b in range(B)])
pr = np.stack([data[b % 16][2] for b in range(B)])
p_r = np.stack([data[b % 16][3] for b in range(B)])
best = 1e9
for _ in range(4):
    T, res, its = api.Levenberg_Marquardt_batch(10, 1e-8, REPS10, [1.0] * 20, 1e-2, T0, pr, p_r, huber_delta=1.0, ctx=ctx)
    best = min(best, ctx.last_kernel_ms())
print("B=%d shape=%s: %.3f ms kernel, mean iters %.1f, r_norm mean %.3e" % (B, os.environ.get("EPIVO_LM_SHAPE", "default"), best, its.mean(), res[:, 1].mean()))
