"""Root-cause analysis of the pairs on which the whole-call census (tests/test_gpu_cv2_census.py) differs from cv2.
TEST TOOLING (uses oracle/ and cv2); run in the build container:  python tests/census_rootcause.py 151:kitti.cpp:101 ...

For each pair: cv2's E / mask, the oracle's (== the GPU's, bit for bit on masks) E / mask, then
  * the Sampson error of every flipped point under both models against the f32 threshold, and
  * cv2's winning model re-scored with the restated rule, the oracle's winner re-scored likewise, and the
    iteration at which each implementation's best model appeared."""
import sys

import numpy as np
import cv2

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from epivo_b200 import synth  # noqa: E402
from oracle import cpu_reference as R  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    seq = synth.make_sequence(257, 2000, seed=synth.seed_for(3, 0))
    Kf = seq.K.astype(np.float32)
    for arg in sys.argv[1:]:
        pair, shape = arg.split(":", 1)
        i = int(pair)
        method, prob, thr = R.CALL_SHAPES[shape]
        qi, ti, _ = O.bf_match(seq.descs[i], seq.descs[i + 1])
        p0, p1 = seq.kps[i][qi], seq.kps[i + 1][ti]
        Ec, mc = cv2.findEssentialMat(p0, p1, Kf, method, prob, thr)
        mc = mc.ravel()
        Eo, mo, info = O.find_essential_mat(p0, p1, Kf, method, prob, thr, 1000)
        x1, x2 = O.normalize_points(p0, Kf), O.normalize_points(p1, Kf)
        t32 = O.ransac_threshold(thr, Kf)
        ec, eo = O.sampson_err_f32(Ec, x1, x2), O.sampson_err_f32(Eo, x1, x2)
        assert np.array_equal(O.find_inliers(ec, t32), mc == 1), "restated rule must reproduce cv2's mask from cv2's E"
        En_c, En_o = Ec / np.linalg.norm(Ec), Eo / np.linalg.norm(Eo)
        dE = min(np.abs(En_c - En_o).max(), np.abs(En_c + En_o).max())
        flips = np.nonzero(mc != mo)[0]
        print(f"pair {i} {shape}: n {len(p0)}  cv2 inliers {int(mc.sum())}  ours {int(mo.sum())}  flipped {len(flips)}  "
              f"|E_cv2 - E_ours| {dE:.2e}  ours iters {info.get('iters')}  thr32 {t32:.9e}")
        if len(flips) <= 8:
            for k in flips:
                print(f"   point {k}: err under cv2's E {ec[k]:.9e}  under ours {eo[k]:.9e}  rel. distance to thr "
                      f"{(ec[k] - t32) / t32:+.2e} / {(eo[k] - t32) / t32:+.2e}")
        else:
            # different winning models: how does each implementation's winner score under the common rule?
            print(f"   different models won.  cv2's E scores {int(O.find_inliers(ec, t32).sum())} inliers under the restated "
                  f"rule, ours {int(O.find_inliers(eo, t32).sum())}")
            # is cv2's winner among our hypotheses?  find the closest of our models over the sample stream
            best = (1e9, None)
            samples = O.generate_samples(len(p0), 1000)
            for it, s in enumerate(samples):
                for E in O.five_point(x1[s], x2[s]):
                    En = E / np.linalg.norm(E)
                    d = min(np.abs(En - En_c).max(), np.abs(En + En_c).max())
                    if d < best[0]:
                        cnt = int(O.find_inliers(O.sampson_err_f32(E, x1, x2), t32).sum())
                        best = (d, it, cnt)
            print(f"   closest of our hypotheses to cv2's winner: sample {best[1]}, |dE| {best[0]:.2e}, our count for it {best[2]}")


if __name__ == "__main__":
    main()
