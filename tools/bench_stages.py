"""Stage-level throughput on one GPU for the configs of BASELINE.json that are not the headline line:
  cfg1 LMedS   the kitti_E.cpp:101 estimator (LMedS, 0.99) through the sequence pipeline
  cfg2         EuRoC-shaped sequence (752x480, 1500 kp, RANSAC 0.99 / 0.3)
  hard RANSAC  kitti_ba.cpp:308 threshold 0.05 (runs to ~1000 iterations)
  cfg4         RANSAC stress: M 5-point hypotheses x N correspondences (K2 samples/s, K3 hypothesis-points/s)
  cfg5         kitti_ba windows: n_zeta = 10, 20 reps x 250 points (windows/s); shipped ws=3 stereo shape
Wall-clock around synchronous C-ABI calls (host copies included where the call takes host buffers)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from epivo_b200 import api, synth

ctx = api.Context(0)
out = {}


def timed(fn, reps=3):
    fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    ctx.sync()
    return (time.perf_counter() - t0) / reps


def seq_rate(name, seq, kp, **kw):
    pipe = api.SequencePipeline(seq.n_frames, kp, ctx=ctx)
    pipe.upload(seq.kps, seq.descs)
    prm = api.default_params(seq.K.astype(np.float32), **kw)
    dt = timed(lambda: pipe.run(prm, 0, seq.n_pairs))
    st = pipe.stage_ms()
    res = pipe.download(0, seq.n_pairs)
    out[name] = {"pairs_per_s": seq.n_pairs / dt, "ms": dt * 1e3, "match_ms": float(st[1]), "essential_ms": float(st[3]),
                 "pose_ms": float(st[4]), "lm_ms": float(st[5]), "mean_iters": float(res["ransac_iters"].mean()),
                 "mean_inliers": float(res["n_inliers"].mean())}
    pipe.close()


F = int(os.environ.get("FRAMES", "1025"))
kitti = synth.make_sequence(F, 2000, seed=synth.seed_for(3, 0))
seq_rate("cfg3_ransac_1.0", kitti, 2000)
seq_rate("cfg1_lmeds", kitti, 2000, method=api.LMEDS, threshold=0.01)
seq_rate("cfg3_ratio0.8_hamming_knn2", kitti, 2000, match_mode=2, norm=api.NORM_HAMMING, ratio=0.8)   # north_star's matcher mode
seq_rate("cfg3_crosscheck_hamming", kitti, 2000, norm=api.NORM_HAMMING)
seq_rate("hard_ransac_0.05", kitti, 2000, threshold=0.05)
euroc = synth.make_sequence(F, 1500, seed=synth.seed_for(2, 0), K=synth.EUROC_K, size=synth.EUROC_SIZE, depth=(1.0, 8.0),
                            px_sigma=0.3, outlier_frac=0.25)
seq_rate("cfg2_euroc_ransac_0.3", euroc, 1500, threshold=0.3)

# cfg4: fixed hypothesis sets -- the whole grid of BASELINE.json config 4 (N x M x outlier fraction)
rng = np.random.default_rng(4)


def knorm(p, K):
    """K-normalised float64 coordinates of float32 pixels (inputs of the stand-alone 5-point solver)."""
    p = np.asarray(p, dtype=np.float32).astype(np.float64)
    return np.stack([(p[:, 0] - K[0, 2]) / K[0, 0], (p[:, 1] - K[1, 2]) / K[1, 1]], axis=1)


GRID = [(N, M, outl) for N in (2000, 8000, 20000) for M in (4096, 16384, 65536) for outl in (0.3, 0.5, 0.7)]
if os.environ.get("CFG4", "grid") == "diag":
    GRID = [(2000, 4096, 0.3), (8000, 16384, 0.5), (20000, 65536, 0.7)]
for N, M, outl in GRID:
    p = synth.make_pair(synth.seed_for(4, N) + int(outl * 10), n=N, outlier_frac=outl)
    Kf = p.K.astype(np.float32)
    # correspondences in row order: true partner for the inliers, a random frame-1 point for the outliers
    c0 = p.kp0
    c1 = np.ascontiguousarray(p.kp1[np.where(p.gt_match >= 0, p.gt_match, rng.integers(0, N, N))])
    x0, x1 = knorm(c0, p.K), knorm(c1, p.K)
    # M samples of 5 distinct indices (argsort of uniform keys: distinct by construction)
    idx = np.argsort(rng.random((M, 64)), axis=1)[:, :5] * (N // 64) + rng.integers(0, N // 64, size=(M, 5))
    X1, X2 = x0[idx], x1[idx]
    dt = timed(lambda: api.fivePointRaw(X1, X2, ctx=ctx), reps=2)
    Es, nm = api.fivePointRaw(X1, X2, ctx=ctx)
    models = np.concatenate([Es[i, :nm[i]] for i in range(M)]).reshape(-1, 9)[: M]
    dt2 = timed(lambda: api.scoreSampson(models, c0, c1, Kf, 1.0, ctx=ctx, medians=False), reps=2)
    cnt = api.scoreSampson(models, c0, c1, Kf, 1.0, ctx=ctx, medians=False)[0]
    out[f"cfg4_N{N}_M{M}_out{int(outl * 100)}"] = {
        "k2_samples_per_s": M / dt, "k3_hyp_points_per_s": len(models) * N / dt2,
        "k3_gflops": len(models) * N * 34 / dt2 / 1e9, "models": int(nm.sum()), "best_inliers": int(np.max(cnt))}

# cfg5: windows
REPS10 = [(i, i) for i in range(10)] + [(0, i) for i in range(10)]
for name, nz, reps, N, B in [("cfg5_nz10_20x250", 10, REPS10, 250, 504),
                             ("kitti_ba_stereo_ws3_9x32", 4, [(0, 1), (1, 1), (0, 0), (0, 3), (1, 3), (0, 0), (2, 3), (3, 3), (2, 2)], 32, 2270)]:
    data = [synth.gen_scene_sequence(500 + b, N, nz, reps) for b in range(min(B, 64))]
    T0 = np.stack([data[b % len(data)][1] for b in range(B)])
    pr = np.stack([data[b % len(data)][2] for b in range(B)])
    p_r = np.stack([data[b % len(data)][3] for b in range(B)])
    dt = timed(lambda: api.Levenberg_Marquardt_batch(nz, 1e-8, reps, [1.0] * len(reps), 1e-2, T0, pr, p_r, huber_delta=1.0, ctx=ctx), reps=2)
    out[name] = {"windows_per_s": B / dt, "ms": dt * 1e3, "windows": B}

print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/stages.json", "w"), indent=1)
