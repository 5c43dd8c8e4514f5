"""TEST TOOLING (GPU box): replay one pair's RANSAC sample stream on the GPU stage by stage and compare with the
oracle -- which sample / model / count first differs.   python tests/gpu_trace_pair.py euroc 213 14"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from epivo_b200 import api, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    kind = sys.argv[1]
    pairs = [int(x) for x in sys.argv[2:]]
    if kind == "euroc":
        seq = synth.make_sequence(max(pairs) + 2, 1500, seed=synth.seed_for(2, 0), K=synth.EUROC_K, size=synth.EUROC_SIZE,
                                  depth=(1.0, 8.0), px_sigma=0.3, outlier_frac=0.25, step=(0.03, 0.07))
        method, prob, thr = 8, 0.99, 0.3
    else:
        seq = synth.make_sequence(max(pairs) + 2, 2000, seed=synth.seed_for(3, 0))
        method, prob, thr = 8, 0.99, float(os.environ.get("THR", "1.0"))
    Kf = seq.K.astype(np.float32)
    ctx = api.Context(0)
    for i in pairs:
        qi, ti, _ = O.bf_match(seq.descs[i], seq.descs[i + 1])
        p0, p1 = seq.kps[i][qi], seq.kps[i + 1][ti]
        n = len(p0)
        Eg, mg, info = api.findEssentialMat(p0, p1, Kf, method, prob, thr, ctx=ctx, return_info=True)
        Eo, mo, oinfo = O.find_essential_mat(p0, p1, Kf, method, prob, thr, 1000)
        print(f"pair {i}: n {n}  GPU inliers {int(mg.sum())} iters {info['iters']}   oracle inliers {int(mo.sum())} iters {oinfo['iters']}"
              f"   masks equal {np.array_equal(mg, mo)}")
        x1, x2 = O.normalize_points(p0, Kf), O.normalize_points(p1, Kf)
        t32 = O.ransac_threshold(thr, Kf)
        S = O.generate_samples(n, max(oinfo["iters"], info["iters"]) + 2)
        Egs, nm = api.fivePointRaw(x1[S], x2[S], ctx=ctx)
        best_o = best_g = 0
        for it, s in enumerate(S):
            Eos = O.five_point(x1[s], x2[s])
            co = [int(O.find_inliers(O.sampson_err_f32(E, x1, x2), t32).sum()) for E in Eos]
            cg = [int(O.find_inliers(O.sampson_err_f32(Egs[it, k], x1, x2), t32).sum()) for k in range(nm[it])]
            same = len(co) == len(cg) and co == cg
            dmax = 0.0
            if len(co) == len(cg):
                for k in range(len(co)):
                    a, b = Eos[k] / np.linalg.norm(Eos[k]), Egs[it, k] / np.linalg.norm(Egs[it, k])
                    dmax = max(dmax, min(np.abs(a - b).max(), np.abs(a + b).max()))
            flag = ""
            if co and max(co) > max(best_o, 4):
                best_o = max(co)
                flag += f" oracle-best {best_o}"
            if cg and max(cg) > max(best_g, 4):
                best_g = max(cg)
                flag += f" gpu-best {best_g}"
            if not same or flag or dmax > 1e-9:
                print(f"   sample {it}: oracle counts {co}  gpu counts {cg}  max |dE| {dmax:.1e}{flag}")
    ctx.close()


if __name__ == "__main__":
    main()
