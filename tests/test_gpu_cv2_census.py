"""Whole-call parity census against LIVE cv2 at benchmark size: for every findEssentialMat call
shape the reference's drivers use, >= 256 synthetic pairs at 2000 keypoints (1500 on the EuRoC camera) go through
cv2 (BFMatcher -> findEssentialMat -> recoverPose, the calls the reference makes) and through the GPU pipeline;
matches, the {0,1} essential mask and the {0,255} pose mask must be identical pair by pair, R / t within
north_star's tolerances.

Known and accepted -- what "the reference's result" can mean here.  OpenCV's five-point solver is not refined: on an
ill-conditioned minimal sample its root violates the essential-matrix constraint 2 E E' E - tr(E E') E = 0 by
1e-6 .. 1e-3 (1e-13 .. 1e-16 otherwise), and which way it errs depends on the null-space basis LAPACK's SVD hands it,
so no independent implementation reproduces that model, its inlier count, or -- through RANSACUpdateNumIters -- the
rest of the trajectory.  On KITTI-shaped pairs (~1 m baseline) this touches 0 .. 1 % of the calls; on EuRoC-shaped
pairs (5 cm baseline at 1 .. 8 m depth: every sample is close to degenerate) ~6 %.  tests/census_rootcause.py and
DESIGN section 2 have the per-pair analysis (with and without the Gauss-Newton polish: without it twice as many
pairs differ).  The rule the test applies:
  * matches, pose masks and counts must agree wherever the essential masks do;
  * a pair whose essential mask differs from cv2's is accepted only if the GPU result is, bit for bit, what the
    oracle (the restatement of OpenCV's algorithm, pinned on cv2 goldens) computes for that pair -- i.e. the
    difference is a property of exact-versus-noisy roots, not of the CUDA code -- or, failing that, if the
    oracle reproduces the GPU's answer once the five points of each hypothesis are perturbed in the 13th digit
    (expanding det B(z) into the degree-10 polynomial loses up to 12 digits on near-degenerate samples -- measured
    against 60-digit arithmetic -- so whether a close pair of roots is real is decided by rounding, in OpenCV too);
  * on the KITTI shapes cv2's own winning E must in addition fail the constraint by > 1e-8 (it always did);
  * the differing fraction is bounded: 2 % (KITTI), 10 % (EuRoC).
The census is written to gpurun_out/cv2_census.json (profiles/r2_cv2_census.json is a committed copy)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N_PAIRS = int(os.environ.get("EPIVO_CENSUS_PAIRS", "256"))
KITTI_SHAPES = ["kitti.cpp:101", "kitti_E.cpp:101", "kitti_ba.cpp:232", "kitti_ba.cpp:308", "kitti_ba.cpp:702"]


def _cubic_residual(E):
    E = E / np.linalg.norm(E)
    return float(np.abs(2 * E @ E.T @ E - np.trace(E @ E.T) * E).max())


def _rot_angle(a, b):
    return float(np.arccos(np.clip((np.trace(a.T @ b) - 1) / 2, -1, 1)))


def _run_census(seq, shapes, tag):
    from epivo_b200 import api
    from oracle import cpu_reference as R
    ref = R.census(seq.kps, seq.descs, seq.K, shapes)
    ctx = api.Context(0)
    pipe = api.SequencePipeline(seq.n_frames, seq.kps.shape[1], ctx=ctx)
    pipe.upload(seq.kps, seq.descs)
    Kf = seq.K.astype(np.float32)
    report = {}
    try:
        for name in shapes:
            method, prob, thr = R.CALL_SHAPES[name]
            prm = api.default_params(Kf, method=method, prob=prob, threshold=thr)
            pipe.run(prm, 0, seq.n_pairs)
            res = pipe.download(0, seq.n_pairs)
            gpu_masks = {}
            rep = {"pairs": seq.n_pairs, "matches_differ": [], "e_mask_differ": [], "pose_mask_differ": [],
                   "n_good_differ": [], "max_rot_diff_rad": 0.0, "max_t_angle_rad": 0.0, "max_E_diff": 0.0}
            for i in range(seq.n_pairs):
                o = ref[i]["shapes"][name]
                qi, ti, _ = pipe.matches(i)
                if not (np.array_equal(qi, ref[i]["qi"]) and np.array_equal(ti, ref[i]["ti"])):
                    rep["matches_differ"].append(i)
                    continue
                if o["E"] is None:
                    continue
                em, pm = pipe.masks(i)
                if not np.array_equal(em, o["e_mask"]):
                    rep["e_mask_differ"].append((i, int((em != o["e_mask"]).sum()) if em.shape == o["e_mask"].shape else -1,
                                                 _cubic_residual(o["E"]), _cubic_residual(res[i]["E"])))
                    gpu_masks[i] = em.copy()
                    continue
                Eg, Ec = res[i]["E"], o["E"]
                Eg, Ec = Eg / np.linalg.norm(Eg), Ec / np.linalg.norm(Ec)
                rep["max_E_diff"] = max(rep["max_E_diff"], float(min(np.abs(Eg - Ec).max(), np.abs(Eg + Ec).max())))
                if not np.array_equal(pm, o["pose_mask"]):
                    rep["pose_mask_differ"].append((i, int((pm != o["pose_mask"]).sum()) if pm.shape == o["pose_mask"].shape else -1))
                    continue
                if int(res[i]["n_good"]) != o["n_good"]:
                    rep["n_good_differ"].append(i)
                rep["max_rot_diff_rad"] = max(rep["max_rot_diff_rad"], _rot_angle(res[i]["R"], o["R"]))
                tg, tc = res[i]["t"], o["t"]
                c = float(tg @ tc / (np.linalg.norm(tg) * np.linalg.norm(tc)))
                rep["max_t_angle_rad"] = max(rep["max_t_angle_rad"], float(np.arccos(np.clip(c, -1, 1))))
            report[name] = rep
            rep["gpu_masks"] = gpu_masks
    finally:
        pipe.close()
        ctx.close()
    os.makedirs("gpurun_out", exist_ok=True)
    path = os.path.join("gpurun_out", "cv2_census.json")
    old = json.load(open(path)) if os.path.exists(path) else {}
    old[tag] = {k: {kk: vv for kk, vv in v.items() if kk != "gpu_masks"} for k, v in report.items()}
    json.dump(old, open(path, "w"), indent=1)
    return report


def _assert_clean(report, seq, max_frac, need_cv2_residual):
    from oracle import cpu_reference as R
    from oracle import oracle as O
    Kf = seq.K.astype(np.float32)
    for name, rep in report.items():
        assert not rep["matches_differ"], (name, rep["matches_differ"][:5])
        assert not rep["pose_mask_differ"], (name, len(rep["pose_mask_differ"]), rep["pose_mask_differ"][:5])
        assert not rep["n_good_differ"], (name, rep["n_good_differ"][:5])
        assert rep["max_rot_diff_rad"] <= 1e-4 and rep["max_t_angle_rad"] <= 1e-3, (name, rep)
        # (pair, differing mask entries, constraint residual of cv2's E, of ours)
        assert len(rep["e_mask_differ"]) <= max(2, int(max_frac * rep["pairs"])), (name, len(rep["e_mask_differ"]))
        method, prob, thr = R.CALL_SHAPES[name]
        for d in rep["e_mask_differ"]:
            i = d[0]
            if need_cv2_residual:
                assert d[2] > 1e-8 and d[3] < 1e-12, (name, d)
            qi, ti, _ = O.bf_match(seq.descs[i], seq.descs[i + 1])
            p0, p1 = seq.kps[i][qi], seq.kps[i + 1][ti]
            Eo, mo, _ = O.find_essential_mat(p0, p1, Kf, method, prob, thr, 1000)
            if np.array_equal(mo, rep["gpu_masks"][i]):
                continue
            # GPU != oracle: accepted only if the pair is a coin flip for ANY float64 implementation -- the oracle's own
            # answer changes when the five sample points of each hypothesis are perturbed in the 13th digit (the
            # expansion of det B(z) into the degree-10 polynomial loses up to 12 digits on such samples, so whether a
            # close pair of roots comes out real is decided by rounding) -- and the GPU's answer is one of the
            # answers the oracle gives under those perturbations.
            outcomes = []
            for t in range(8):
                rng = np.random.default_rng(1000 * i + t)

                def solver(a, b, rng=rng):
                    return O.five_point(a * (1 + 1e-13 * rng.standard_normal(a.shape)),
                                        b * (1 + 1e-13 * rng.standard_normal(b.shape)))
                outcomes.append(O.find_essential_mat(p0, p1, Kf, method, prob, thr, 1000, solver=solver)[1])
            assert any(np.array_equal(m, rep["gpu_masks"][i]) for m in outcomes), \
                (name, "GPU != oracle on pair", i, "and no 1e-13 perturbation of the samples reproduces the GPU's answer")
            rep.setdefault("rounding_decided", []).append(i)


def test_census_kitti_shape_2000kp_all_kitti_call_sites():
    from epivo_b200 import synth
    seq = synth.make_sequence(N_PAIRS + 1, 2000, seed=synth.seed_for(3, 0))
    _assert_clean(_run_census(seq, KITTI_SHAPES, "kitti_2000kp"), seq, 0.02, True)


def test_census_euroc_shape_1500kp():
    from epivo_b200 import synth
    seq = synth.make_sequence(N_PAIRS + 1, 1500, seed=synth.seed_for(2, 0), K=synth.EUROC_K, size=synth.EUROC_SIZE,
                              depth=(1.0, 8.0), px_sigma=0.3, outlier_frac=0.25, step=(0.03, 0.07))
    _assert_clean(_run_census(seq, ["euroc_E.cpp:205"], "euroc_1500kp"), seq, 0.10, False)
