import numpy as np, sys, os
sys.path.insert(0,'/root/repo')
from epivo_b200 import api, synth
kind=sys.argv[1]
if kind=='euroc':
    seq = synth.make_sequence(513, 1500, seed=synth.seed_for(2, 0), K=synth.EUROC_K, size=synth.EUROC_SIZE, depth=(1.0, 8.0), px_sigma=0.3, outlier_frac=0.25, step=(0.03, 0.07)); kw=dict(method=8,prob=.99,threshold=.3)
else:
    seq = synth.make_sequence(513, 2000, seed=synth.seed_for(3, 0)); kw=dict(method=int(sys.argv[2]),prob=.99,threshold=float(sys.argv[3]))
ctx=api.Context(0); pipe=api.SequencePipeline(seq.n_frames, seq.kps.shape[1], ctx=ctx); pipe.upload(seq.kps,seq.descs)
prm=api.default_params(seq.K.astype(np.float32), **kw)
pipe.run(prm,0,seq.n_pairs); ctx.sync()
print(pipe.stage_ms()[:8])
