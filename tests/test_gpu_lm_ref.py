"""GPU LM (epivo_lm_rt / epivo_lm_rt_batch through the C ABI) against the reference's OWN Levenberg_Marquardt:
tests/golden/lm_ref.npz holds the outputs of /root/reference/jac_Rt_gen_.cpp compiled unmodified (oracle/_ref,
tests/golden/make_golden_lm_ref.py) for the shapes the drivers and the demo run -- kitti_E.cpp:196 (1 zeta, 48
points), euroc_E.cpp:283-299, the demo's 10-zeta chain (test_jac_Rt_gen.cpp:282-297) on scenes drawn by the
reference's own generator, BASELINE config 5 (20 reps x 250), the kitti_ba stereo window (ws = 3), reverse reps,
a w = 0 rep, a singular H.  Comparison rule and tolerances: tests/lm_ref_util.py."""
import numpy as np
import pytest

from lm_ref_util import DELTAS, EPSILON, LAMBDA0, NAMES, case, check, gold

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from epivo_b200 import api
    c = api.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("tag", ["ref", "d1"])
@pytest.mark.parametrize("name", NAMES)
def test_gpu_lm_equals_reference_build(ctx, name, tag):
    from epivo_b200 import api
    c = case(name)
    T, info = api.Levenberg_Marquardt(c["n_zeta"], EPSILON, c["reps"], c["wreps"], LAMBDA0, c["T0s"], c["pr"], c["p_r"],
                                      huber_delta=DELTAS[tag], ctx=ctx)
    check(name, tag, T, info["r_norm"], info["lambda"], info["H_norm"])
    if gold(name, tag)["nan_break"]:
        assert np.array_equal(T, c["T0s"])                    # "delta has Nan" before any update (:407)


@pytest.mark.parametrize("shape", ["384x128x2", "256x64x4", "192x96", "128x64"])
def test_gpu_lm_cfg5_every_kernel_shape_equals_reference_build(ctx, shape, monkeypatch):
    """The window kernel's CTA / cluster shapes differ in summation order only: each one must meet the
    reference on BASELINE config 5 (and on the demo chain) at both Huber settings."""
    from epivo_b200 import api
    monkeypatch.setenv("EPIVO_LM_SHAPE", shape)
    for name in ("cfg5_51", "demo_refgen_3"):
        c = case(name)
        for tag in ("ref", "d1"):
            T, info = api.Levenberg_Marquardt(c["n_zeta"], EPSILON, c["reps"], c["wreps"], LAMBDA0, c["T0s"], c["pr"],
                                              c["p_r"], huber_delta=DELTAS[tag], ctx=ctx)
            check(name, tag, T, info["r_norm"], info["lambda"], info["H_norm"])


def test_gpu_lm_batch_of_kitti_e_pairs_equals_reference_build(ctx):
    """The single-pair kernel (8 lanes per problem) on a batch: the three kitti_E goldens side by side."""
    from epivo_b200 import api
    names = [n for n in NAMES if n.startswith("kitti_E")]
    cs = [case(n) for n in names]
    for tag in ("ref", "d1"):
        Tb, res, its = api.Levenberg_Marquardt_batch(1, EPSILON, [(0, 0)], [1.0], LAMBDA0, np.stack([c["T0s"] for c in cs]),
                                                     np.stack([c["pr"] for c in cs]), np.stack([c["p_r"] for c in cs]),
                                                     huber_delta=DELTAS[tag], ctx=ctx)
        for k, n in enumerate(names):
            check(n, tag, Tb[k], res[k][1], res[k][2], res[k][0])
