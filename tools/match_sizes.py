"""GPU tuning / reporting helper: the matcher at the reference's other descriptor-set sizes -- 1500 x 1500 (EuRoC shape,
BASELINE config 2) and 10000 x 10000 (ORB::create(10000), kitti_ba.cpp:128) next to the headline 2000 x 2000.
Prints, per size: matcher tile-kernel ms per pair, descriptor pairs/s, and the issued POPC.32 rate as a fraction of the
POPC peak measured in the same run (epivo_microbench 0).  JSON on stdout (profiles/r2_matcher_sizes.json)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api, synth

ctx = api.Context(0)
popc_peak = ctx.microbench(0)
out = {"popc32_peak_Gops": popc_peak / 1e9, "sizes": []}
for kp, frames in ((1500, 2049), (2000, 2049), (10000, 97)):
    seq = synth.make_sequence(frames, kp, seed=synth.seed_for(3, 1))
    pipe = api.SequencePipeline(seq.n_frames, kp, ctx=ctx)
    pipe.upload(seq.kps, seq.descs)
    prm = api.default_params(seq.K.astype(np.float32))
    for _ in range(3):
        pipe.run(prm, 0, seq.n_pairs)
    ctx.sync()
    ms = pipe.stage_ms()
    res = pipe.download(0, seq.n_pairs)
    per_pair_us = ms[7] / seq.n_pairs * 1e3
    alg_popc = kp * kp * 4 * seq.n_pairs / (ms[7] * 1e-3)           # Hamming2 on bit planes: 4 words per descriptor pair
    out["sizes"].append({"kp": kp, "pairs": seq.n_pairs, "match_kernel_ms": float(ms[7]), "us_per_pair": float(per_pair_us),
                         "pairs_per_s_matcher_only": float(seq.n_pairs / (ms[7] * 1e-3)),
                         "descriptor_pairs_per_s": float(kp * kp * seq.n_pairs / (ms[7] * 1e-3)),
                         "issued_popc_frac_of_peak": float(alg_popc * 0.75 / popc_peak),
                         "whole_pipeline_pairs_per_s": float(seq.n_pairs / (ms[0] * 1e-3)),
                         "mean_matches": float(res["n_matches"].mean())})
    pipe.close()
print(json.dumps(out, indent=1))
