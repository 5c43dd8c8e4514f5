"""Comparison rule for the LK tracker (floating point: the tolerance is stated here, once, for the oracle and the GPU).

cv2 sums its 2 x 2 system in float32 SSE lanes; the restatement and the GPU sum the same integer products exactly
and round once.  Both are faithful evaluations of the same formula, so positions agree to ~1e-4 px -- except when a
stopping test of the iteration (|delta|^2 <= 1e-4, the oscillation test, the minimum-eigenvalue gate) is decided
the other way by that last bit, which moves a point by at most one LK step (< 0.1 px at convergence) or flips its
status.  Rule: status equal on >= 99.5 % of the points; among the points both track, >= 99 % within 1e-3 px and all
within 0.25 px."""
import numpy as np


def check_err(err, st, ref_err, ref_st, what=""):
    """OpenCV's `err` (mean absolute window difference at the final position, a sum of integers scaled once): where both
    track a point it follows the position, so it agrees to the same 1e-3 relative to the 0..255 grey range."""
    both = (np.asarray(st).ravel() == 1) & (np.asarray(ref_st).ravel() == 1)
    if both.any():
        d = np.abs(np.asarray(err).ravel()[both] - np.asarray(ref_err).ravel()[both])
        assert (d <= 1e-3).mean() >= 0.99 and d.max() <= 0.5, (what, float(d.max()))


def check_lk(nxt, st, ref_nxt, ref_st, what=""):
    nxt, ref_nxt = np.asarray(nxt).reshape(-1, 2), np.asarray(ref_nxt).reshape(-1, 2)
    st, ref_st = np.asarray(st).ravel(), np.asarray(ref_st).ravel()
    assert len(st) == len(ref_st), what
    if len(st) == 0:
        return 0.0
    same = st == ref_st
    assert same.mean() >= 0.995, (what, "status differs on", int((~same).sum()), "of", len(st))
    both = (st == 1) & (ref_st == 1)
    if not both.any():
        return 0.0
    d = np.abs(nxt[both] - ref_nxt[both]).max(axis=1)
    assert (d <= 1e-3).mean() >= 0.99, (what, "more than 1 % beyond 1e-3 px", float(np.percentile(d, 99)))
    assert d.max() <= 0.25, (what, float(d.max()))
    return float(d.max())
