"""Generates tests/golden/lm_ref.npz: inputs and outputs of the reference's OWN Levenberg_Marquardt
(/root/reference/jac_Rt_gen_.cpp:287-478, compiled unmodified into oracle/_ref by oracle/Makefile),
for the shapes the reference's drivers and demo run.  Run in the build container (needs /root/reference):

    python tests/golden/make_golden_lm_ref.py

Every case is run twice: `ref` = the unmodified build (huber_delta = 1e-5, jac_Rt_gen_.cpp:17) and
`d1` = the build with that one constant patched to 1.0 (test_jac_Rt_gen.cpp:16).  Stored per case:
n_zeta, reps, wreps, T0s, pr, p_r (inputs) and, per build, T (refined chain), H_norm, r_norm, lambda,
accepts, rejects (from lambda = lambda0 * 5^rej / 2^acc), nan_break, trace (|delta| and candidate |r0| per
iteration, read off the stand-in's norm() calls).  The reference LM always runs <= 30 iterations with
epsilon / lambda0 as given (1e-8 / 1e-2 at every call site).

Conditioning.  The accept test `curr_E < prev_E` (jac_Rt_gen_.cpp:457) compares, once the iteration has
converged on noisy data, two norms that agree to ~1e-14, and as lambda -> 0 the damped H is nearly singular
along the monocular scale gauge; some scenes (reverse reps, whose Jacobian is the reference's left-perturbation
quirk) amplify a 1e-13 difference by 10x per iteration.  What an independent implementation -- or the same
source against real Eigen -- can reproduce is therefore case-dependent, and is MEASURED here: every case is
re-run `N_PERTURB` times with all inputs multiplied by (1 + u * 2^-52), u in {-1, 0, 1}.  Stored per build:
`stable` (accept / reject / nan sequence identical in all runs), `sens` = [max relative change of r_norm,
max relative change of H_norm, max absolute change of T].  tests/lm_ref_util.py turns these into the
comparison rule: exact step counts where `stable`, r_norm within max(1e-5, 20 * sens).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from epivo_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import pipeline as OP  # noqa: E402
from oracle import reflib  # noqa: E402

REPS10 = [r for i in range(10) for r in ((i, i), (0, i))]           # test_jac_Rt_gen.cpp:294-297
N_PERTURB = 4
STEREO_WS3 = [(0, 1), (1, 1), (0, 0), (0, 3), (1, 3), (0, 0), (2, 3), (3, 3), (2, 2)]   # kitti_ba.cpp:931-938,969-975


def kitti_e_inputs(index, method, thr, make=synth.make_kitti_pair, n=2000):
    """LM inputs exactly as kitti_E.cpp:137-191 builds them from the pair's E / recoverPose result."""
    pr_ = make(index, n)
    Kf = pr_.K.astype(np.float32)
    out = OP.pair_pipeline(pr_.kp0, pr_.desc0, pr_.kp1, pr_.desc1, Kf, method=method, thr=thr, lm_iters=0)
    qi, ti, _ = out["matches"]
    m = out["e_mask"] == 1
    c0, c1 = pr_.kp0[qi][m], pr_.kp1[ti][m]
    x0, x1 = O.normalize_points(c0[:48], Kf), O.normalize_points(c1[:48], Kf)
    pr = np.concatenate([x0, np.ones((48, 1))], axis=1)[None]
    p_r = np.concatenate([x1, np.ones((48, 1))], axis=1)[None]
    return out["T0"][None].copy(), pr, p_r


def main():
    assert reflib.build(), "oracle/_ref could not be built (is /root/reference present?)"
    R, D1 = reflib.ref(), reflib.ref_d1()
    assert R.huber_delta == 1e-5 and D1.huber_delta == 1.0
    cases = {}

    def add(name, n_zeta, reps, w, T0s, pr, p_r):
        cases[name] = dict(n_zeta=n_zeta, reps=np.array(reps, np.int32), wreps=np.array(w, np.float64),
                           T0s=np.array(T0s), pr=np.array(pr), p_r=np.array(p_r))

    # kitti_E.cpp:196 shape: one zeta, one rep, 48 points, T0 from recoverPose
    for k, (meth, thr) in enumerate(((O.LMEDS, 0.01), (O.RANSAC, 1.0), (O.LMEDS, 0.01))):
        T0s, pr, p_r = kitti_e_inputs(k, meth, thr)
        add(f"kitti_E_{k}", 1, [(0, 0)], [1.0], T0s, pr, p_r)
    T0s, pr, p_r = kitti_e_inputs(0, O.RANSAC, 0.3, synth.make_euroc_pair, 1500)      # euroc_E.cpp:283-299
    add("euroc_E_0", 1, [(0, 0)], [1.0], T0s, pr, p_r)
    # the demo's shape, scenes drawn by the reference's own generator (sequence.hpp:106-159, srand(seed))
    for seed in (3, 11):
        Ts, T0s, Xr, pr, p_r = R.gen_scene_sequence(seed, 15, 10, REPS10)
        add(f"demo_refgen_{seed}", 10, REPS10, [1.0] * 20, T0s, pr, p_r)
    # BASELINE config 5: 10 zetas, 20 reps x 250 points
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(51, 250, 10, REPS10)
    add("cfg5_51", 10, REPS10, [1.0] * 20, T0s, pr, p_r)
    # kitti_ba.cpp stereo window (ws = 3): 9 reps x 32 points over 4 zetas; then one '<32 points' rep
    # (weight 0, all-ones dummy points, kitti_ba.cpp:983-987)
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(61, 32, 4, STEREO_WS3)
    add("stereo_ws3", 4, STEREO_WS3, [1.0] * 9, T0s, pr, p_r)
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(63, 32, 4, STEREO_WS3)
    pr[4][:] = 1.0
    p_r[4][:] = 1.0
    w = [1.0] * 9
    w[4] = 0.0
    add("stereo_ws3_w0", 4, STEREO_WS3, w, T0s, pr, p_r)
    # reverse reps (RepJacobian's inverse branch, jac_Rt_gen_.cpp:276-281)
    reps = [(0, 0), (1, 1), (2, 2), (3, 3), (0, 3), (3, 1), (2, 0), (1, 0)]
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(71, 24, 4, reps)
    add("reverse_reps", 4, reps, [1.0, 1.0, 1.0, 1.0, 0.5, 2.0, 1.0, 0.25], T0s, pr, p_r)
    # a zeta observed only by a w = 0 rep: singular H, "delta has Nan", break (jac_Rt_gen_.cpp:407)
    reps = [(0, 0), (1, 1)]
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(5, 32, 2, reps)
    pr[1][:] = 1.0
    p_r[1][:] = 1.0
    add("w0_singular", 2, reps, [1.0, 0.0], T0s, pr, p_r)

    rng = np.random.default_rng(2024)

    def ulp(a):
        return a * (1.0 + rng.integers(-1, 2, size=a.shape) * 2.0 ** -52)

    flat = {}
    for name, c in cases.items():
        for k, v in c.items():
            flat[f"{name}/{k}"] = v
        for tag, lib in (("ref", R), ("d1", D1)):
            T, info = lib.levenberg_marquardt(c["n_zeta"], 1e-8, c["reps"], c["wreps"], 1e-2, c["T0s"], c["pr"], c["p_r"])
            flat[f"{name}/{tag}/T"] = T
            flat[f"{name}/{tag}/scalars"] = np.array([info["H_norm"], info["r_norm"], info["lambda"]])
            flat[f"{name}/{tag}/steps"] = np.array([info["accepts"], info["rejects"], int(info["nan_break"])], np.int32)
            flat[f"{name}/{tag}/trace"] = np.array([(d, np.nan if e is None else e) for d, e in info["trace"]]).reshape(-1, 2)
            stable, sr, sh, st = True, 0.0, 0.0, 0.0
            for _ in range(N_PERTURB):
                T2, i2 = lib.levenberg_marquardt(c["n_zeta"], 1e-8, c["reps"], c["wreps"], 1e-2, ulp(c["T0s"]),
                                                 ulp(c["pr"]), ulp(c["p_r"]))
                stable &= (i2["accepts"], i2["rejects"], i2["nan_break"]) == (info["accepts"], info["rejects"], info["nan_break"])
                sr = max(sr, abs(i2["r_norm"] - info["r_norm"]) / info["r_norm"])
                sh = max(sh, abs(i2["H_norm"] - info["H_norm"]) / info["H_norm"])
                st = max(st, float(np.abs(T2 - T).max()))
            flat[f"{name}/{tag}/stable"] = np.array(stable)
            flat[f"{name}/{tag}/sens"] = np.array([sr, sh, st])
            print(f"{name:18s} {tag:3s} iters {info['iters']:2d} (acc {info['accepts']:2d}) nan {int(info['nan_break'])} "
                  f"r_norm {info['r_norm']:.6e} lambda {info['lambda']:.3e}  stable {stable}  sens r {sr:.1e} H {sh:.1e} T {st:.1e}")
    flat["names"] = np.array(sorted(cases))
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lm_ref.npz")
    np.savez_compressed(out, **flat)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
