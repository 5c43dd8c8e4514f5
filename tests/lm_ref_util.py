"""Shared by the CPU and GPU LM parity tests: access to tests/golden/lm_ref.npz (outputs of the reference's own
jac_Rt_gen_.cpp, see tests/golden/make_golden_lm_ref.py) and the comparison rule.

north_star's tolerances: relative residual <= 1e-5, rotation <= 1e-4 rad, unit-t angle <= 1e-3 rad.  They are
applied as stated wherever the reference itself is reproducible under a 1-ulp perturbation of its inputs; where
the golden file records that it is not (`stable` False: the accept / reject sequence after convergence is decided
by rounding; `sens`: measured change of the outputs under that perturbation), the step counts are not compared
and the tolerance is widened to 20 x the measured sensitivity."""
import math
import os

import numpy as np

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lm_ref.npz"))
NAMES = [str(n) for n in G["names"]]
DELTAS = {"ref": 1e-5, "d1": 1.0}      # ref: unmodified build; d1: huber_delta patched to 1.0 (test_jac_Rt_gen.cpp:16)
LAMBDA0, EPSILON = 1e-2, 1e-8          # every reference call site (kitti_E.cpp:196, kitti_ba.cpp:881)


def case(name):
    c = {k: G[f"{name}/{k}"] for k in ("reps", "wreps", "T0s", "pr", "p_r")}
    c["n_zeta"] = int(G[f"{name}/n_zeta"])
    return c


def gold(name, tag):
    H, r, lam = G[f"{name}/{tag}/scalars"]
    acc, rej, nanb = (int(v) for v in G[f"{name}/{tag}/steps"])
    sr, sh, st = G[f"{name}/{tag}/sens"]
    return dict(T=G[f"{name}/{tag}/T"], H_norm=H, r_norm=r, lam=lam, acc=acc, rej=rej, nan_break=bool(nanb),
                stable=bool(G[f"{name}/{tag}/stable"]), sens_r=sr, sens_H=sh, sens_T=st,
                trace=G[f"{name}/{tag}/trace"])


def lambda_steps(lam, lambda0=LAMBDA0, max_iters=64):
    """(accepts, rejects) with lam == lambda0 * 5**rejects / 2**accepts (unique: 2 and 5 are coprime)."""
    found = []
    for rej in range(max_iters + 1):
        acc = int(round(math.log2(lambda0 * 5.0 ** rej / lam)))
        if 0 <= acc <= max_iters - rej and abs(lambda0 * 5.0 ** rej / 2.0 ** acc - lam) <= 1e-9 * lam:
            found.append((acc, rej))
    assert len(found) == 1, (lam, found)
    return found[0]


def rot_angle(Ra, Rb):
    c = (np.trace(Ra.T @ Rb) - 1.0) / 2.0
    return math.acos(max(-1.0, min(1.0, c))) if c < 1 - 1e-12 else float(np.linalg.norm(Ra - Rb) / math.sqrt(2))


def check(name, tag, T, r_norm, lam, H_norm=None):
    """Compare one implementation's result for case `name` with the reference build `tag`."""
    g = gold(name, tag)
    T = np.asarray(T)
    if g["stable"]:
        assert lambda_steps(lam) == (g["acc"], g["rej"]), (name, tag, lambda_steps(lam), g["acc"], g["rej"])
        assert abs(lam - g["lam"]) <= 1e-12 * g["lam"]
    tol_r = max(1e-5, 20 * g["sens_r"])
    assert abs(r_norm - g["r_norm"]) <= tol_r * g["r_norm"] + 1e-16, (name, tag, r_norm, g["r_norm"], tol_r)
    if H_norm is not None and g["stable"]:
        assert abs(H_norm - g["H_norm"]) <= max(1e-5, 20 * g["sens_H"]) * g["H_norm"] + 1e-16, (name, tag, H_norm, g["H_norm"])
    n_zeta = T.shape[0]
    for k in range(n_zeta):
        assert rot_angle(T[k][:3, :3], g["T"][k][:3, :3]) <= max(1e-4, 20 * g["sens_T"]), (name, tag, k)
    if n_zeta == 1:
        # one pair: the length of t is the unobservable monocular scale (it drifts as lambda -> 0); direction only
        a, b = T[0][:3, 3], g["T"][0][:3, 3]
        cosang = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))
        assert math.acos(min(1.0, cosang)) <= 1e-3, (name, tag)
    else:
        tol_t = max(1e-6 * max(1.0, float(np.abs(g["T"][:, :3, 3]).max())), 20 * g["sens_T"])
        assert np.abs(T[:, :3, 3] - g["T"][:, :3, 3]).max() <= tol_t, (name, tag, np.abs(T[:, :3, 3] - g["T"][:, :3, 3]).max())


def check_trace(name, tag, trace, tie=1e-11):
    """Step-by-step comparison for implementations that expose their trajectory: same accept / reject decision
    and candidate residual (relative 1e-6) at every iteration before the reference's first rounding-level
    decision (|curr_E - prev_E| <= tie * prev_E, or |delta| within 1 % of epsilon)."""
    g = gold(name, tag)
    prev_g = prev = 1e10
    n = 0
    for i, (dn_g, e_g) in enumerate(g["trace"]):
        if abs(dn_g - EPSILON) <= 1e-2 * EPSILON or i >= len(trace):
            break
        dn, e = trace[i]
        if e_g != e_g:                       # the reference broke here (NaN or |delta| < epsilon)
            assert e is None, (name, tag, i)
            break
        if abs(e_g - prev_g) <= tie * prev_g:
            break
        assert e is not None and (e < prev) == (e_g < prev_g), (name, tag, i, e, prev, e_g, prev_g)
        if g["sens_r"] < 1e-7:
            assert abs(e - e_g) <= 1e-6 * e_g, (name, tag, i, e, e_g)
        if e_g < prev_g:
            prev_g, prev = e_g, e
        n += 1
    return n
