"""CPU restatement of the detector the reference's front end calls (TEST INFRASTRUCTURE: imported only by tests/,
__graft_entry__.smoke() and bench.py's CPU legs -- never by the product path).

cv::FastFeatureDetector (FAST-9/16) is an OpenCV dependency of the reference, not reference source:
    Ptr<FastFeatureDetector> detector = FastFeatureDetector::create(40);  detector->detect(src, kp0, Mat());
        kitti_E.cpp:71-74, kitti_ba.cpp:49,62 (threshold 40), kitti_ba.cpp:98,117-118 (default threshold 10).
OpenCV is un-vendored and un-versioned in the reference (compile_cv: `pkg-config opencv`); the algorithm restated here
is the published one of modules/features2d/src/fast.cpp (FAST_t<16>) and fast_score.cpp (cornerScore<16>), and it is
pinned against `cv2 4.13.0` -- live in tests/test_oracle_fast.py and through tests/golden/fast.npz.
"""
from __future__ import annotations

import numpy as np

# (dx, dy) of the 16-pixel circle in OpenCV's order (fast_score.cpp makeOffsets, patternSize 16)
CIRCLE = [(0, 3), (1, 3), (2, 2), (3, 1), (3, 0), (3, -1), (2, -2), (1, -3),
          (0, -3), (-1, -3), (-2, -2), (-3, -1), (-3, 0), (-3, 1), (-2, 2), (-1, 3)]


def _arc9(mask: np.ndarray) -> np.ndarray:
    """mask: (16, h, w) bool -> (h, w) bool, true where 9 circularly contiguous entries are set."""
    out = np.zeros(mask.shape[1:], dtype=bool)
    for s in range(16):
        run = np.ones(mask.shape[1:], dtype=bool)
        for k in range(9):
            run &= mask[(s + k) & 15]
        out |= run
    return out


def fast_scores(img: np.ndarray, threshold: int):
    """(corner (rows, cols) bool, score (rows, cols) int) of FAST_t<16>: fast.cpp -- a pixel with a full circle
    (3 <= x < cols-3, 3 <= y < rows-3) is a corner when 9 contiguous circle pixels are all < v - t or all > v + t
    (the pairwise rejection cascade in front of that test never rejects one); score = cornerScore<16>."""
    img = np.asarray(img, dtype=np.uint8)
    rows, cols = img.shape
    corner = np.zeros((rows, cols), dtype=bool)
    score = np.zeros((rows, cols), dtype=np.int32)
    threshold = int(threshold)
    if not 0 <= threshold <= 255:      # beyond 8 bits cv2's vector path truncates the threshold and its scalar tail does not
        raise ValueError("threshold outside [0, 255]")
    if rows < 7 or cols < 7:
        return corner, score
    v = img[3:rows - 3, 3:cols - 3].astype(np.int32)
    d = np.stack([v - img[3 + dy:rows - 3 + dy, 3 + dx:cols - 3 + dx].astype(np.int32) for dx, dy in CIRCLE])
    c = _arc9(d > threshold) | _arc9(-d > threshold)
    # cornerScore<16>: a0 = max(t, max over the 16 arcs of 9 of min d); b0 = min(-a0, min over arcs of max d); -b0-1
    a0 = np.full(v.shape, threshold, dtype=np.int32)
    arcs_min, arcs_max = [], []
    for s in range(16):
        idx = [(s + k) & 15 for k in range(9)]
        arcs_min.append(d[idx].min(axis=0))
        arcs_max.append(d[idx].max(axis=0))
    a0 = np.maximum(a0, np.max(arcs_min, axis=0))
    b0 = np.minimum(-a0, np.min(arcs_max, axis=0))
    sc = -b0 - 1
    corner[3:rows - 3, 3:cols - 3] = c
    score[3:rows - 3, 3:cols - 3] = np.where(c, sc, 0)
    return corner, score


def fast_detect(img: np.ndarray, threshold: int = 10, nonmax: bool = True):
    """cv2.FastFeatureDetector_create(threshold, nonmax).detect(img): (pts (k, 2) float32 [x, y], response (k,) float32)
    in OpenCV's order (row by row, left to right).  With suppression a corner survives when its score is strictly
    greater than the scores of its 8 neighbours (0 for non-corners); without it the response is 0 (fast.cpp keeps the
    zero-initialised score row)."""
    corner, score = fast_scores(img, threshold)
    if nonmax:
        s = np.pad(score, 1)
        keep = corner.copy()
        rows, cols = score.shape
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                if dx or dy:
                    keep &= score > s[1 + dy:1 + dy + rows, 1 + dx:1 + dx + cols]
    else:
        keep = corner
    ys, xs = np.nonzero(keep)                                   # row-major: OpenCV's order
    pts = np.stack([xs, ys], axis=1).astype(np.float32).reshape(-1, 2)
    resp = (score[ys, xs] if nonmax else np.zeros(len(xs))).astype(np.float32)
    return pts, resp


# ---------------------------------------------------------------------------------------------------------------
# calcOpticalFlowPyrLK (OpenCV video/src/lkpyramid.cpp), as the reference calls it with all defaults
#     calcOpticalFlowPyrLK(src, tgt, pt0, pt1_, status, err);          kitti_E.cpp:79-84, kitti_ba.cpp:203-208,281-286
# i.e. winSize 21 x 21, maxLevel 3, criteria (COUNT + EPS, 30, 0.01), flags 0, minEigThreshold 1e-4.
# Integer parts (pyramid, Scharr derivatives, the fixed-point bilinear patches) are exact; the float parts follow
# OpenCV's float32 formulas, but its 2 x 2 system is accumulated in float32 in the lane order of its SSE code, which
# this restatement (sums in exact integers, then one rounding) does not imitate: results agree with cv2 to ~1e-3 px,
# except where a stopping test falls on the other side (tests/test_oracle_lk.py states the tolerances).

def _reflect101(i, n):
    i = np.abs(i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def pyr_down(img: np.ndarray) -> np.ndarray:
    """cv::pyrDown for 8-bit images: separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8, size (n + 1) / 2."""
    img = np.asarray(img, dtype=np.uint8)
    rows, cols = img.shape
    orow, ocol = (rows + 1) // 2, (cols + 1) // 2
    w = np.array([1, 4, 6, 4, 1], dtype=np.int64)
    xi = _reflect101(2 * np.arange(ocol)[:, None] + np.arange(-2, 3)[None, :], cols)       # (ocol, 5)
    yi = _reflect101(2 * np.arange(orow)[:, None] + np.arange(-2, 3)[None, :], rows)
    h = (img.astype(np.int64)[:, xi] * w).sum(axis=2)                                       # (rows, ocol)
    v = (h[yi, :] * w[None, :, None]).sum(axis=1)                                           # (orow, ocol)
    return ((v + 128) >> 8).astype(np.uint8)


def scharr_deriv(img: np.ndarray) -> np.ndarray:
    """calcSharrDeriv: (rows, cols, 2) int16 [dI/dx, dI/dy], 3-10-3 Scharr, BORDER_REFLECT_101."""
    img = np.asarray(img, dtype=np.uint8).astype(np.int32)
    rows, cols = img.shape
    ym, yp = _reflect101(np.arange(rows) - 1, rows), _reflect101(np.arange(rows) + 1, rows)
    t0 = (img[ym] + img[yp]) * 3 + img * 10
    t1 = img[yp] - img[ym]
    xm, xp = _reflect101(np.arange(cols) - 1, cols), _reflect101(np.arange(cols) + 1, cols)
    dx = t0[:, xp] - t0[:, xm]
    dy = (t1[:, xp] + t1[:, xm]) * 3 + t1 * 10
    return np.stack([dx, dy], axis=2).astype(np.int16)


def build_pyramid(img: np.ndarray, win: int = 21, max_level: int = 3):
    """buildOpticalFlowPyramid: levels while both sides stay larger than the window."""
    pyr = [np.asarray(img, dtype=np.uint8)]
    for _ in range(max_level):
        rows, cols = pyr[-1].shape
        if (cols + 1) // 2 <= win or (rows + 1) // 2 <= win:
            break
        pyr.append(pyr_down(pyr[-1]))
    return pyr


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


def _weights(a, b):
    f = np.float32
    iw00 = np.rint((f(1) - a) * (f(1) - b) * f(1 << 14)).astype(np.int64)
    iw01 = np.rint(a * (f(1) - b) * f(1 << 14)).astype(np.int64)
    iw10 = np.rint((f(1) - a) * b * f(1 << 14)).astype(np.int64)
    return iw00, iw01, iw10, (1 << 14) - iw00 - iw01 - iw10


def _patch(img, ix, iy, wts, win, shift, zero_outside):
    """Bilinear fixed-point window of every point: (P, win, win) int64.  img (rows, cols) integer; outside the image
    BORDER_REFLECT_101 (the pyramid's border) or 0 (the derivative's BORDER_CONSTANT)."""
    rows, cols = img.shape
    ys = iy[:, None] + np.arange(win + 1)[None, :]
    xs = ix[:, None] + np.arange(win + 1)[None, :]
    if zero_outside:
        oky = (ys >= 0) & (ys < rows)
        okx = (xs >= 0) & (xs < cols)
        g = img[np.clip(ys, 0, rows - 1)[:, :, None], np.clip(xs, 0, cols - 1)[:, None, :]].astype(np.int64)
        g = g * (oky[:, :, None] & okx[:, None, :])
    else:
        g = img[_reflect101(ys, rows)[:, :, None], _reflect101(xs, cols)[:, None, :]].astype(np.int64)
    iw00, iw01, iw10, iw11 = (w[:, None, None] for w in wts)
    v = g[:, :-1, :-1] * iw00 + g[:, :-1, 1:] * iw01 + g[:, 1:, :-1] * iw10 + g[:, 1:, 1:] * iw11
    return _descale(v, shift)


def calc_optical_flow_pyr_lk(prev: np.ndarray, nxt: np.ndarray, pts: np.ndarray, win: int = 21, max_level: int = 3,
                             max_count: int = 30, epsilon: float = 0.01, min_eig_threshold: float = 1e-4,
                             with_err: bool = True, return_err: bool = False):
    """(next_pts (P, 2) float32, status (P,) uint8 [, err (P,) float32]) of cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None)
    with the reference's defaults.  Follows LKTrackerInvoker::operator() level by level; all point arithmetic in
    float32.  with_err: the caller passes an `err` output (every reference call site and cv2 do): OpenCV then measures
    the window difference at the final position and clears the status if that position is out of range (the iteration
    loop itself does not re-check the position its last update produced)."""
    f = np.float32
    pts = np.asarray(pts, dtype=np.float32).reshape(-1, 2)
    P = len(pts)
    pp, pn = build_pyramid(prev, win, max_level), build_pyramid(nxt, win, max_level)
    L = min(len(pp), len(pn)) - 1
    status = np.ones(P, dtype=np.uint8)
    errv = np.zeros(P, dtype=np.float32)
    nextp = np.zeros((P, 2), dtype=np.float32)
    half = f((win - 1) * 0.5)
    eps2 = f(epsilon) * f(epsilon)          # criteria.epsilon *= criteria.epsilon (a double in OpenCV; compared with a double dot)
    eps2 = float(epsilon) * float(epsilon)
    FLT_SCALE = f(1.0 / (1 << 20))
    for level in range(L, -1, -1):
        I, J = pp[level], pn[level]
        rows, cols = I.shape
        dI = scharr_deriv(I)
        prevPt = pts * f(1.0 / (1 << level))
        nextPt = prevPt.copy() if level == L else nextp * f(2)
        nextp = nextPt.copy()
        prevPt = prevPt - half
        ip = np.floor(prevPt).astype(np.int64)
        oob = (ip[:, 0] < -win) | (ip[:, 0] >= cols) | (ip[:, 1] < -win) | (ip[:, 1] >= rows)
        if level == 0:
            status[oob] = 0
        act = ~oob
        if not act.any():
            continue
        idx = np.nonzero(act)[0]
        a = prevPt[idx, 0] - ip[idx, 0].astype(f)
        b = prevPt[idx, 1] - ip[idx, 1].astype(f)
        wts = _weights(a, b)
        Iw = _patch(I, ip[idx, 0], ip[idx, 1], wts, win, 14 - 5, False)
        Ix = _patch(dI[:, :, 0], ip[idx, 0], ip[idx, 1], wts, win, 14, True)
        Iy = _patch(dI[:, :, 1], ip[idx, 0], ip[idx, 1], wts, win, 14, True)
        A11 = (Ix * Ix).sum(axis=(1, 2)).astype(f) * FLT_SCALE
        A12 = (Ix * Iy).sum(axis=(1, 2)).astype(f) * FLT_SCALE
        A22 = (Iy * Iy).sum(axis=(1, 2)).astype(f) * FLT_SCALE
        D = A11 * A22 - A12 * A12
        minEig = (A22 + A11 - np.sqrt((A11 - A22) * (A11 - A22) + f(4) * A12 * A12)) / f(2 * win * win)
        bad = (minEig < f(min_eig_threshold)) | (D < f(np.finfo(np.float32).eps))
        if level == 0:
            status[idx[bad]] = 0
        keep = ~bad
        idx, Iw, Ix, Iy = idx[keep], Iw[keep], Ix[keep], Iy[keep]
        A11, A12, A22, D = A11[keep], A12[keep], A22[keep], D[keep]
        with np.errstate(divide="ignore"):
            D = f(1) / D
        npt = nextPt[idx] - half
        prevDelta = np.zeros((len(idx), 2), dtype=f)
        live = np.ones(len(idx), dtype=bool)
        for j in range(max_count):
            if not live.any():
                break
            li = np.nonzero(live)[0]
            inp = np.floor(npt[li]).astype(np.int64)
            out = (inp[:, 0] < -win) | (inp[:, 0] >= cols) | (inp[:, 1] < -win) | (inp[:, 1] >= rows)
            if level == 0:
                status[idx[li[out]]] = 0
            live[li[out]] = False
            li, inp = li[~out], inp[~out]
            if len(li) == 0:
                break
            a = npt[li, 0] - inp[:, 0].astype(f)
            b = npt[li, 1] - inp[:, 1].astype(f)
            Jw = _patch(J, inp[:, 0], inp[:, 1], _weights(a, b), win, 14 - 5, False)
            diff = Jw - Iw[li]
            b1 = (diff * Ix[li]).sum(axis=(1, 2)).astype(f) * FLT_SCALE
            b2 = (diff * Iy[li]).sum(axis=(1, 2)).astype(f) * FLT_SCALE
            delta = np.stack([(A12[li] * b2 - A22[li] * b1) * D[li], (A12[li] * b1 - A11[li] * b2) * D[li]], axis=1).astype(f)
            npt[li] = npt[li] + delta
            nextp[idx[li]] = npt[li] + half
            dd = delta[:, 0].astype(np.float64) ** 2 + delta[:, 1].astype(np.float64) ** 2
            conv = dd <= eps2
            osc = np.zeros(len(li), dtype=bool)
            if j > 0:
                osc = (~conv) & (np.abs(delta[:, 0] + prevDelta[li, 0]) < 0.01) & (np.abs(delta[:, 1] + prevDelta[li, 1]) < 0.01)
                nextp[idx[li[osc]]] = nextp[idx[li[osc]]] - delta[osc] * f(0.5)
            live[li[conv | osc]] = False
            prevDelta[li] = delta
        if level == 0 and with_err:
            # err: mean absolute window difference at the final nextPt - halfWin (same floor / fixed-point weights as an
            # iteration); a final position outside [-win, size) clears the status
            ok = np.nonzero(status[idx] == 1)[0]
            np_ = nextp[idx[ok]] - half
            ir = np.floor(np_).astype(np.int64)
            out = (ir[:, 0] < -win) | (ir[:, 0] >= cols) | (ir[:, 1] < -win) | (ir[:, 1] >= rows)
            status[idx[ok[out]]] = 0
            ok, np_, ir = ok[~out], np_[~out], ir[~out]
            if len(ok):
                aa = np_[:, 0] - ir[:, 0].astype(f)
                bb = np_[:, 1] - ir[:, 1].astype(f)
                Jw = _patch(J, ir[:, 0], ir[:, 1], _weights(aa, bb), win, 14 - 5, False)
                errv[idx[ok]] = np.abs(Jw - Iw[ok]).sum(axis=(1, 2)).astype(f) * f(1.0 / (32 * win * win))
    if return_err:
        return nextp, status, errv
    return nextp, status


# ---------------------------------------------------------------------------------------------------------------
# cv::remap with the fixed-point maps of initUndistortRectifyMap (OpenCV imgproc: imgwarp.cpp remapBilinear,
# undistort.dispatch.cpp), as the EuRoC driver uses them:
#     initUndistortRectifyMap(cam, dist, rect, proj, Size(w, h), map1.type(), map1, map2);     euroc_E.cpp:105-113
#     remap(src, src_, map1, map2, INTER_LINEAR);                                              euroc_E.cpp:169-174
# (`map1.type()` of the still-empty Mat is 0, which selects CV_16SC2 + CV_16UC1).

def remap_bilinear_fixed(img: np.ndarray, map_xy: np.ndarray, map_frac: np.ndarray, border: int = 0) -> np.ndarray:
    """cv2.remap(img, map_xy, map_frac, INTER_LINEAR), BORDER_CONSTANT: integer position + 5 + 5 fraction bits,
    weights {(32-fx)(32-fy), fx(32-fy), (32-fx)fy, fx fy} * 32, (sum + 2^14) >> 15; source pixels outside the image
    are the border value."""
    img = np.asarray(img, dtype=np.uint8)
    rows, cols = img.shape
    sx = map_xy[:, :, 0].astype(np.int64)
    sy = map_xy[:, :, 1].astype(np.int64)
    f = map_frac.astype(np.int64) & 1023
    fx, fy = f & 31, f >> 5
    pad = np.full((rows + 2, cols + 2), border, dtype=np.int64)          # one border ring: everything further out is constant too
    pad[1:-1, 1:-1] = img

    def px(y, x):
        return pad[np.clip(y + 1, 0, rows + 1), np.clip(x + 1, 0, cols + 1)]
    v = (px(sy, sx) * (32 - fx) * (32 - fy) + px(sy, sx + 1) * fx * (32 - fy) + px(sy + 1, sx) * (32 - fx) * fy +
         px(sy + 1, sx + 1) * fx * fy) * 32
    return np.clip((v + (1 << 14)) >> 15, 0, 255).astype(np.uint8)


def init_undistort_rectify_map(K, dist, R, P, size):
    """initUndistortRectifyMap(K, dist, R, P, (w, h), CV_16SC2): (map_xy (h, w, 2) int16, map_frac (h, w) uint16) for the
    radial-tangential model k1 k2 p1 p2 [k3] (what the EuRoC calibration has).  Evaluated directly in float64; OpenCV
    walks every row incrementally (x += ir[0] ...), so a position that falls on a 1/32-pixel rounding boundary can come
    out one step apart (tests/test_oracle_remap.py: at most a handful of pixels per 752 x 480 map)."""
    K = np.asarray(K, dtype=np.float64).reshape(3, 3)
    d = np.zeros(5)
    d[:len(np.ravel(dist))] = np.ravel(dist)[:5]
    k1, k2, p1, p2, k3 = d
    R = np.eye(3) if R is None else np.asarray(R, dtype=np.float64).reshape(3, 3)
    P = np.asarray(P, dtype=np.float64)
    iR = np.linalg.inv(P[:3, :3] @ R)
    w, h = size
    u, v = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
    X = iR[0, 0] * u + iR[0, 1] * v + iR[0, 2]
    Y = iR[1, 0] * u + iR[1, 1] * v + iR[1, 2]
    W = iR[2, 0] * u + iR[2, 1] * v + iR[2, 2]
    x, y = X / W, Y / W
    x2, y2 = x * x, y * y
    r2 = x2 + y2
    _2xy = 2 * x * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
    mu = K[0, 0] * xd + K[0, 2]
    mv = K[1, 1] * yd + K[1, 2]
    iu = np.rint(mu * 32).astype(np.int64)                  # saturate_cast<int>(u * INTER_TAB_SIZE)
    iv = np.rint(mv * 32).astype(np.int64)
    xy = np.stack([np.clip(iu >> 5, -32768, 32767), np.clip(iv >> 5, -32768, 32767)], axis=2).astype(np.int16)
    frac = ((iv & 31) * 32 + (iu & 31)).astype(np.uint16)
    return xy, frac
