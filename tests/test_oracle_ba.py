"""CPU: window bookkeeping of the batched BA orchestration against the sequential restatement
(no LM: the LM function is replaced by a deterministic stand-in, so this runs without a GPU)."""
import numpy as np

from epivo_b200 import ba, synth
from oracle import ba as OBA


def fake_lm(n_zeta, eps, reps, wreps, lam, T0s, pr, p_r, huber_delta=1e-5):
    """Deterministic stand-in: nudges every pose, reports a residual that sometimes trips the revert rule."""
    T0s = np.array(T0s, dtype=np.float64)
    out = T0s.copy()
    out[:, :3, 3] *= 1.0 + 0.01 * np.arange(1, n_zeta + 1)[:, None]
    key = float(np.abs(np.asarray(pr)).sum() + np.asarray(wreps).sum())
    r_norm = 2e-2 if int(key * 1e3) % 3 == 0 else 1e-3
    return out, {"H_norm": key, "r_norm": r_norm, "lambda": lam, "iters": 1}


def _batched_with(fake, reprojs, window, stride, F, K, stereo):
    win = ba.expand_stereo_window(window) if stereo else list(window)
    nodes, step = (2 * F, 2) if stereo else (F, 1)
    starts = ba.window_starts(win, stride, nodes, step)
    opt = np.tile(np.eye(4), (nodes, 1, 1))
    T0s, pr, p_r, w, nz, lo = ba.assemble(reprojs, win, starts, K, stereo)
    reps = ba.window_reps(win)
    T_opt, lm = [], []
    for b in range(len(starts)):
        To, l = fake(nz, 1e-8, reps, w[b], 1e-2, T0s[b], pr[b], p_r[b])
        T_opt.append(To)
        lm.append([l["H_norm"], l["r_norm"], l["lambda"]])
    return (*ba.finish(opt, T0s, np.array(T_opt), np.array(lm), starts, nz, lo, step, stereo), starts)


def test_window_bookkeeping_matches_sequential_loop():
    K = synth.KITTI_K
    for stereo, ws, stride, F in [(False, 3, 2, 15), (False, 4, 1, 11), (True, 3, 2, 9)]:
        window = []
        for i in range(ws - 1):
            window.append((i, i + 1))
            if i < ws - 2:
                window.append((i, i + 2))
        reprojs = synth.make_reprojs(80 + ws, F, window, stereo=stereo, few_points_at={(2, 4)})
        opt, lm, rev, starts = _batched_with(fake_lm, reprojs, window, stride, F, K, stereo)
        o_opt, o_lm, o_rev, o_starts = OBA.bundle_adjustment(reprojs, window, stride, F, K, stereo, lm=fake_lm)
        assert starts == o_starts and len(starts) > 0
        assert np.array_equal(rev, o_rev) and rev.any() and not rev.all()
        assert np.allclose(opt, o_opt, atol=1e-14) and np.allclose(lm, o_lm)


def test_stereo_window_expansion_and_reps():
    """kitti_ba.cpp:934-941 and :969-975 for ws = 3: the nine reps of tests/test_gpu_pose_lm.py."""
    win = ba.expand_stereo_window([(0, 1), (0, 2), (1, 2)])
    assert win == [(0, 2), (1, 2), (0, 1), (0, 4), (1, 4), (0, 1), (2, 4), (3, 4), (2, 3)]
    assert ba.window_reps(win) == [(0, 1), (1, 1), (0, 0), (0, 3), (1, 3), (0, 0), (2, 3), (3, 3), (2, 2)]


def test_window_pairs_order_and_oracle_match_kp():
    """The keys match_kp visits (kitti_ba.cpp:603-615): window offsets per start frame, repeats skipped, the inner
    loop broken at the first pair past the end; and the CPU restatement fills exactly those keys."""
    from epivo_b200 import ba, synth
    from oracle import ba as OBA
    window = [(0, 1), (0, 2), (1, 2)]
    assert ba.window_pairs(window, 2, 5) == [(0, 1), (0, 2), (1, 2), (2, 3), (2, 4), (3, 4)]
    assert ba.window_pairs(window, 1, 4) == [(0, 1), (0, 2), (1, 2), (1, 3), (2, 3)]       # (3,4) breaks the inner loop
    assert ba.window_pairs(window, 2, 1) == []
    seq = synth.make_sequence(n_frames=3, n=300, seed=synth.seed_for(3, 90))
    rp = OBA.match_kp(seq.kps, seq.descs, window, 2, seq.K)
    assert list(rp.keys()) == ba.window_pairs(window, 2, 3)
    p0, p1, R, t = rp[(0, 1)]
    assert len(p0) == len(p1) > 50 and abs(np.linalg.det(R) - 1) < 1e-9 and abs(np.linalg.norm(t) - 1) < 1e-9
