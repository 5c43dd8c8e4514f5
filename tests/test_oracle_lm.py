"""The LM restatement (jac_Rt_gen_.cpp) has no golden vectors upstream (its demos seed with time(0));
it is pinned by finite differences of `res`, by the independent single-pair closed form of
deprecated/test_jac_Rt.cpp, and by the convergence criterion the reference's demo prints."""
import numpy as np

from epivo_b200 import synth
from oracle import oracle as O

REPS = [(i, i) for i in range(4)] + [(0, i) for i in range(4)] + [(3, 1)]


def _fd(T0s, zeta, src, tgt, p, p_, left):
    J = np.zeros((p.shape[0], 6))
    h = 1e-7
    for k in range(6):
        rr = []
        for sgn in (+1, -1):
            T2 = list(T0s)
            d = np.zeros(6)
            d[k] = sgn * h
            T2[zeta] = (O.se3_exp(d) @ T0s[zeta]) if left else (T0s[zeta] @ O.se3_exp(d))
            T = synth.compose_rep(T2, src, tgt)
            rr.append(O.res(T[:3, :3], T[:3, 3], p, p_, 1.0))
        J[:, k] = (rr[0] - rr[1]) / (2 * h)
    return J


def test_forward_jacobian_matches_finite_differences():
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(7, 15, 4, REPS)
    mem = O.chain_memo(list(T0s))
    for zeta, src, tgt, j in ((1, 0, 3, 7), (0, 0, 0, 0), (3, 0, 3, 7)):
        J = O.rep_jacobian(mem, zeta, src, tgt, pr[j], p_r[j], 1.0)
        assert np.abs(J - _fd(T0s, zeta, src, tgt, pr[j], p_r[j], left=False)).max() < 5e-9


def test_reverse_jacobian_is_a_left_perturbation_quirk():
    """Reference quirk (jac_Rt_gen_.cpp:276-281): for reverse reps the code differentiates
    T_zeta -> exp(eps) T_zeta although the update right-multiplies.  Restated as written."""
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(7, 15, 4, REPS)
    mem = O.chain_memo(list(T0s))
    for zeta in (1, 2, 3):
        J = O.rep_jacobian(mem, zeta, 3, 1, pr[8], p_r[8], 1.0)
        assert np.abs(J - _fd(T0s, zeta, 3, 1, pr[8], p_r[8], left=True)).max() < 5e-9
    J = O.rep_jacobian(mem, 2, 3, 1, pr[8], p_r[8], 1.0)
    assert np.abs(J - _fd(T0s, 2, 3, 1, pr[8], p_r[8], left=False)).max() > 1e-3


def test_huber_branch_quirk_sqrt2():
    """jac_Rt_gen_.cpp:203-207 vs :255-257: in the Huber region the Jacobian is sqrt(2) x grad(res)."""
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(9, 20, 1, [(0, 0)])
    d = 1e-5
    T = T0s[0]
    raw = O.res(T[:3, :3], T[:3, 3], pr[0], p_r[0], 1e30)          # e'e / 2 before the Huber switch
    huber = raw > d                                                # res' branch (jac_Rt_gen_.cpp:255)
    assert huber.sum() >= 10 and (2 * raw[huber] > d).all()        # ... all also in the Jacobian's branch (:203)
    J = O.dr_deps(T, np.eye(4), pr[0], p_r[0], False, d)
    Jfd = np.zeros_like(J)
    h = 1e-7
    for k in range(6):
        e = np.zeros(6)
        e[k] = h
        Tp, Tm = T @ O.se3_exp(e), T @ O.se3_exp(-e)
        Jfd[:, k] = (O.res(Tp[:3, :3], Tp[:3, 3], pr[0], p_r[0], d) - O.res(Tm[:3, :3], Tm[:3, 3], pr[0], p_r[0], d)) / (2 * h)
    for i in range(J.shape[0]):
        big = np.abs(Jfd[i]) > 1e-7
        if not big.any() or (not huber[i] and 2 * raw[i] > d):
            continue                                               # between the two switches: mixed regime
        want = np.sqrt(2.0) if huber[i] else 1.0
        assert np.allclose(J[i][big] / Jfd[i][big], want, rtol=1e-4)


def test_closed_form_single_pair():
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(3, 25, 1, [(0, 0)])
    T = T0s[0]
    a = O.dr_deps(T, np.eye(4), pr[0], p_r[0], False, 1.0)
    b = O.single_pair_jacobian_closed_form(T[:3, :3], T[:3, 3], pr[0], p_r[0])
    assert np.abs(a - b).max() < 1e-12


def test_lm_converges_like_the_reference_demo():
    """test_jac_Rt_gen.cpp:464-509 prints |R - R0| and t/t0 per zeta: rotation recovered, translation
    recovered up to one common scale."""
    reps = [(i, i) for i in range(3)] + [(0, i) for i in range(3)]
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(5, 15, 3, reps)
    Tout, info = O.levenberg_marquardt(3, 1e-8, reps, [1.0] * len(reps), 1e-2, T0s, pr, p_r, huber_delta=1.0,
                                       max_iters=60)
    assert info["r_norm"] < 1e-10
    ratios = []
    for k in range(3):
        assert np.linalg.norm(Tout[k][:3, :3] - Ts[k][:3, :3]) < 1e-6
        ratios.append(Ts[k][:3, 3] / Tout[k][:3, 3])
    assert np.ptp(np.concatenate(ratios)) < 1e-5


def test_lm_shipped_delta_does_not_converge():
    """With huber_delta = 1e-5 as shipped (jac_Rt_gen_.cpp:17) the same start stalls: lambda grows."""
    Ts, T0s, pr, p_r = synth.gen_scene_sequence(5, 15, 1, [(0, 0)])
    Tout, info = O.levenberg_marquardt(1, 1e-8, [(0, 0)], [1.0], 1e-2, T0s, pr, p_r)
    assert info["iters"] == 30 and info["lambda"] > 1.0 and info["r_norm"] > 1e-7


def test_se3_exp_small_angle_and_group_property():
    d = np.array([0.1, -0.2, 0.3, 0.01, 0.02, -0.03])
    T = O.se3_exp(d)
    assert np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3)).max() < 1e-14
    assert np.abs(O.se3_exp(d) @ O.se3_exp(-d) - np.eye(4)).max() < 1e-14
    z = O.se3_exp(np.array([1.0, 2.0, 3.0, 0, 0, 0]))
    assert np.allclose(z[:3, 3], [1, 2, 3]) and np.allclose(z[:3, :3], np.eye(3))


def test_c_restatement_matches_numpy_restatement():
    from oracle import clib
    cases = [(1, [(0, 0)], 48, 1e-5, 30), (1, [(0, 0)], 48, 1.0, 30),
             (3, [(0, 0), (1, 1), (2, 2), (0, 2), (2, 0), (1, 0)], 20, 1.0, 20),
             (3, [(0, 0), (1, 1), (2, 2), (0, 2), (2, 0), (1, 0)], 20, 1e-5, 20)]
    for n_zeta, reps, N, delta, iters in cases:
        Ts, T0s, pr, p_r = synth.gen_scene_sequence(40 + n_zeta, N, n_zeta, reps)
        w = [1.0] * len(reps)
        Ta, a = O.levenberg_marquardt(n_zeta, 1e-8, reps, w, 1e-2, T0s, pr, p_r, huber_delta=delta, max_iters=iters)
        Tb, b = clib.levenberg_marquardt(n_zeta, 1e-8, reps, w, 1e-2, T0s, pr, p_r, huber_delta=delta, max_iters=iters)
        assert a["iters"] == b["iters"]
        assert np.abs(Ta - Tb).max() < 1e-7
        assert abs(a["r_norm"] - b["r_norm"]) <= 1e-6 * a["r_norm"] + 1e-18
        assert abs(a["lambda"] - b["lambda"]) <= 1e-12 * a["lambda"]
