"""Distinct contexts are fully concurrent (the reference runs association and BA on different std::threads,
kitti_ba.cpp:1153-1163; its LM keeps a mutable global, jac_Rt_gen_.cpp:20 -- this library has none)."""
import threading

import numpy as np
import pytest

from epivo_b200 import api, synth

pytestmark = pytest.mark.gpu


def _work(ctx, seed, out, key):
    p = synth.make_kitti_pair(seed, n=800)
    Kf = p.K.astype(np.float32)
    qi, ti, _ = api.BFMatcher(api.NORM_HAMMING2, True, ctx=ctx).match(p.desc0, p.desc1)
    r = api.findEssentialMat(p.kp0[qi], p.kp1[ti], Kf, api.RANSAC, 0.99, 1.0, ctx=ctx)
    E, mask = r[0], r[1]
    n, R, t, pm = api.recoverPose(E, p.kp0[qi][mask == 1], p.kp1[ti][mask == 1], Kf, ctx=ctx)
    reps = [(0, 0), (1, 1), (0, 1)]
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(seed, 40, 2, reps)
    T, lm = api.Levenberg_Marquardt(2, 1e-8, reps, [1.0] * 3, 1e-2, T0s, pr, p_r, huber_delta=1.0, ctx=ctx)
    out[key] = (qi, ti, E, mask, n, R, t, pm, T, lm["r_norm"])


def test_two_contexts_on_two_threads_match_serial_results():
    ctxs = [api.Context(0), api.Context(0)]
    serial, conc = {}, {}
    for k in range(2):
        _work(ctxs[k], 30 + k, serial, k)
    for rep in range(3):
        th = [threading.Thread(target=_work, args=(ctxs[k], 30 + k, conc, k)) for k in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for k in range(2):
            for a, b in zip(serial[k], conc[k]):
                assert np.array_equal(np.asarray(a), np.asarray(b))
    for c in ctxs:
        c.close()
