import numpy as np, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epivo_b200 import api, synth
from oracle import oracle as O
i, samp = int(sys.argv[1]), int(sys.argv[2])
seq = synth.make_sequence(i + 2, 1500, seed=synth.seed_for(2, 0), K=synth.EUROC_K, size=synth.EUROC_SIZE, depth=(1.0, 8.0), px_sigma=0.3, outlier_frac=0.25, step=(0.03, 0.07))
Kf = seq.K.astype(np.float32)
qi, ti, _ = O.bf_match(seq.descs[i], seq.descs[i + 1]); p0, p1 = seq.kps[i][qi], seq.kps[i + 1][ti]
x1, x2 = O.normalize_points(p0, Kf), O.normalize_points(p1, Kf)
s = O.generate_samples(len(p0), samp + 1)[samp]
ctx = api.Context(0)
E, nm = api.fivePointRaw(x1[s][None], x2[s][None], ctx=ctx)
ctx.sync()
print("gpu models", nm)
Eo = O.five_point(x1[s], x2[s])
print("oracle models", len(Eo))
for k in range(nm[0]):
    d = [min(np.abs(E[0, k] - e).max(), np.abs(E[0, k] + e).max()) for e in Eo]
    print(" gpu model", k, "closest oracle", int(np.argmin(d)), "%.1e" % min(d))
