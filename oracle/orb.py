"""CPU restatement of cv::ORB::detectAndCompute as the reference configures it (TEST INFRASTRUCTURE: imported only by
tests/, __graft_entry__.smoke() and bench.py's CPU legs -- never by the product path).

    Ptr<ORB> orb = ORB::create(10000, 1.2f, 8, 15, 0, 2, ORB::FAST_SCORE);     kitti_ba.cpp:128
    orb->detect(src, kp);  orb->compute(src, kp, desc);                          kitti_ba.cpp:131-152
(nfeatures 10000, scaleFactor 1.2, nlevels 8, edgeThreshold 15, firstLevel 0, WTA_K 2, FAST score, and OpenCV's
defaults patchSize 31, fastThreshold 20).  OpenCV is an un-vendored, un-versioned dependency of the reference
(compile_cv: `pkg-config opencv`); the algorithm restated here is the published one of modules/features2d/src/orb.cpp
(ORB_Impl::detectAndCompute, computeKeyPoints, ICAngles, computeOrbDescriptors), imgproc's bit-exact resize
(INTER_LINEAR_EXACT) and fixed-point GaussianBlur, and KeyPointsFilter::retainBest, pinned against `cv2 4.13.0` -- live in
tests/test_oracle_orb.py and through tests/golden/orb.npz.  The 256 x 4 sampling pattern (`bit_pattern_31_`) is data of
that dependency: epivo_b200/orb_pattern.py holds it, tools/extract_orb_pattern.py reads it out of the installed cv2.
"""
from __future__ import annotations

import math

import numpy as np

from . import frontend as OF

HARRIS_BLOCK_SIZE = 9
PATCH_SIZE = 31
HALF_PATCH = PATCH_SIZE // 2


def _cv_round(x) -> int:
    """cvRound: round half to even (SSE cvtsd2si / cvtss2si)."""
    return int(np.rint(x))


def layer_scales(nlevels: int, scale_factor: float, first_level: int = 0) -> np.ndarray:
    """orb.cpp getScale: (float)pow((double)scaleFactor, level - firstLevel); scaleFactor is the float the caller passed."""
    sf = float(np.float32(scale_factor))
    return np.array([np.float32(math.pow(sf, lv - first_level)) for lv in range(nlevels)], dtype=np.float32)


def layer_sizes(rows: int, cols: int, scales: np.ndarray):
    """orb.cpp detectAndCompute:  float inv_scale = 1.f / scale;  Size sz(cvRound(cols * inv_scale), cvRound(rows * inv_scale))
    -- the product with the float reciprocal, not a division: 285 columns at scale 1.2f give 237.5 -> 238 this way and
    237.49998 -> 237 by division (cv2 4.13.0 returns the 238-column level; tests/orb_census.py found the case)."""
    out = []
    for s in scales:
        inv = np.float32(np.float32(1.0) / np.float32(s))
        out.append((_cv_round(np.float32(np.float32(rows) * inv)), _cv_round(np.float32(np.float32(cols) * inv))))
    return out


def features_per_level(nfeatures: int, nlevels: int, scale_factor: float) -> list[int]:
    """computeKeyPoints: geometric split of nfeatures over the levels (float arithmetic as in orb.cpp)."""
    factor = np.float32(1.0 / float(np.float32(scale_factor)))           # (float)(1.0 / scaleFactor)
    # float ndesired = nfeatures*(1 - factor)/(1 - (float)pow((double)factor, (double)nlevels));  all in float
    nd = np.float32(np.float32(np.float32(nfeatures) * np.float32(np.float32(1) - factor)) /
                    np.float32(np.float32(1) - np.float32(math.pow(float(factor), float(nlevels)))))
    out, total = [], 0
    for _ in range(nlevels - 1):
        n = _cv_round(nd)
        out.append(n)
        total += n
        nd = np.float32(nd * factor)
    out.append(max(nfeatures - total, 0))
    return out


def _lin_coeffs(src: int, dst: int):
    """resize.cpp interpolationLinear<ufixedpoint16>::getCoeffs for every destination index: source offset, the two
    8.8 fixed-point weights, and the [minofst, maxofst) range outside which the edge pixel is copied."""
    inv_scale = np.float64(dst) / np.float64(src)
    scale = np.float64(1.0) / inv_scale
    ofs = np.zeros(dst, np.int64)
    c1 = np.zeros(dst, np.int64)
    lo, hi = 0, dst
    for v in range(dst):
        fval = scale * (np.float64(v) + 0.5) - 0.5
        iv = int(np.floor(fval))
        if iv >= 0 and src > 1:
            if iv < src - 1:
                ofs[v] = iv
                c1[v] = int(np.rint((fval - iv) * 256.0))
            else:
                ofs[v] = src - 1
                hi = min(hi, v)
        else:
            lo = max(lo, v + 1)
    return ofs, 256 - c1, c1, lo, hi


def resize_linear_exact(img: np.ndarray, drows: int, dcols: int) -> np.ndarray:
    """cv::resize(src, dst, Size(dcols, drows), 0, 0, INTER_LINEAR_EXACT), 8-bit: horizontal pass into 8.8 fixed point,
    vertical pass into 16.16, rounded once."""
    srows, scols = img.shape
    ox, x0, x1, xlo, xhi = _lin_coeffs(scols, dcols)
    oy, y0, y1, ylo, yhi = _lin_coeffs(srows, drows)
    im = img.astype(np.int64)
    H = np.empty((srows, dcols), np.int64)
    H[:, :xlo] = im[:, :1] * 256
    xs = np.arange(xlo, xhi)
    H[:, xs] = im[:, ox[xs]] * x0[xs] + im[:, np.minimum(ox[xs] + 1, scols - 1)] * x1[xs]
    H[:, xhi:] = im[:, scols - 1:scols] * 256
    ys = np.arange(drows)
    top = np.clip(oy, 0, srows - 1)
    V = H[top] * y0[:, None] + H[np.minimum(top + 1, srows - 1)] * y1[:, None]
    V[ys < ylo] = H[0] * 256
    V[ys >= yhi] = H[srows - 1] * 256
    return np.clip((V + (1 << 15)) >> 16, 0, 255).astype(np.uint8)


# getGaussianKernel(7, 2.0, CV_32F): exp(-x^2 / 8) normalised in double, stored as float32
_g = np.exp(-np.arange(-3, 4, dtype=np.float64) ** 2 / 8.0)
GAUSS7 = (_g / _g.sum()).astype(np.float32)


def _fma32(k, x, acc):
    """float32 fused multiply-add on arrays: the product of two float32 is exact in float64; the float64 sum is rounded
    to float32 afterwards (differs from a true FMA only when that sum falls within 2^-29 of a float32 tie)."""
    return (np.float64(k) * x.astype(np.float64) + acc.astype(np.float64)).astype(np.float32)


def gaussian_blur_7x7(ext: np.ndarray, border: int) -> np.ndarray:
    """What `GaussianBlur(workingMat, workingMat, Size(7, 7), 2, 2, BORDER_REFLECT_101)` does to a pyramid level inside
    ORB: workingMat is a ROI of the level-with-margin buffer and the border type is not ISOLATED, so OpenCV leaves its
    bit-exact fixed-point path (GaussianBlur: `!src.isSubmatrix()`) for sepFilter2D with the float32 kernel, reading the
    real neighbours (the reflected margin) around the ROI.  Arithmetic of filter.simd.hpp as the AVX2 build runs it:
    rows   RowFilter<uchar, float>:        s = k[0]*p[0];  s = fma(k[i], p[i], s)           i = 1..6, left to right
    cols   SymmColumnFilter<float, uchar>: s = k[3]*r[0];  s = fma(k[3+i], r[i] + r[-i], s)  i = 1..3;  cvRound, saturate
    (pinned by counting descriptor bits against cv2 over 42 k keypoints: the other orders / unfused forms lose 1-3
    descriptors; a fused or unfused row pass is not distinguishable on that sample).  Returns the buffer with its
    interior blurred and the margin untouched -- descriptors sample both."""
    rows, cols = ext.shape[0] - 2 * border, ext.shape[1] - 2 * border
    p = ext[border - 3:border + rows + 3, border - 3:border + cols + 3].astype(np.float32)
    H = (GAUSS7[0] * p[:, 0:cols]).astype(np.float32)
    for i in range(1, 7):
        H = _fma32(GAUSS7[i], p[:, i:i + cols], H)
    V = (GAUSS7[3] * H[3:3 + rows]).astype(np.float32)
    for i in (1, 2, 3):
        V = _fma32(GAUSS7[3 + i], (H[3 + i:3 + i + rows] + H[3 - i:3 - i + rows]).astype(np.float32), V)
    out = ext.copy()
    out[border:border + rows, border:border + cols] = np.clip(np.rint(V), 0, 255).astype(np.uint8)
    return out


def build_pyramid(img: np.ndarray, nlevels: int = 8, scale_factor: float = 1.2, edge_threshold: int = 15):
    """The level images of ORB_Impl::detectAndCompute: level 0 is the image, level k the INTER_LINEAR_EXACT resize of
    level k-1; each is held with a BORDER_REFLECT_101 margin of `border` pixels.  Returns (scales, [ext images], border)."""
    desc_patch = int(math.ceil(HALF_PATCH * math.sqrt(2.0)))
    border = max(edge_threshold, desc_patch, HARRIS_BLOCK_SIZE // 2) + 1
    scales = layer_scales(nlevels, scale_factor)
    sizes = layer_sizes(img.shape[0], img.shape[1], scales)
    exts, prev = [], img
    for lv, (r, c) in enumerate(sizes):
        cur = img if lv == 0 else resize_linear_exact(prev, r, c)
        exts.append(np.pad(cur, border, mode="reflect"))
        prev = cur
    return scales, exts, border


def umax_table(half_patch: int = HALF_PATCH) -> list[int]:
    """computeKeyPoints: the circular patch's half-width per row."""
    umax = [0] * (half_patch + 2)
    vmax = int(math.floor(half_patch * math.sqrt(2.0) / 2 + 1))
    vmin = int(math.ceil(half_patch * math.sqrt(2.0) / 2))
    for v in range(vmax + 1):
        umax[v] = _cv_round(math.sqrt(float(half_patch * half_patch - v * v)))
    v0 = 0
    for v in range(half_patch, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return umax[:half_patch + 1]


_DEG = np.float32(180 / math.pi)                       # static const float atan2_pK = <float literal>*(float)(180/CV_PI)
_ATAN_P1 = np.float32(np.float32(0.9997878412794807) * _DEG)
_ATAN_P3 = np.float32(np.float32(-0.3258083974640975) * _DEG)
_ATAN_P5 = np.float32(np.float32(0.1555786518463281) * _DEG)
_ATAN_P7 = np.float32(np.float32(-0.04432655554792128) * _DEG)
_F_EPS = np.float32(2.220446049250313e-16)


def _fma(a, b, c):
    """float32 fused multiply-add: the product of two float32 is exact in float64; the float64 sum is then rounded to
    float32 (a double rounding that differs from a true FMA only on exact ties of the float64 sum -- not reachable with
    the 24-bit operands here in any case the tests have met)."""
    return np.float32(np.float64(a) * np.float64(b) + np.float64(c))


def fast_atan2(y, x, fused: bool = False):
    """cv::fastAtan2(float y, float x) in degrees: mathfuncs_core atan_f32's degree-7 odd polynomial.  cv2 4.13.0
    evaluates the Horner steps unfused (every angle of 42 k keypoints equal); `fused` is the contracted form, kept to
    show the difference (1 ulp on ~1.5 % of the angles)."""
    x, y = np.float32(x), np.float32(y)
    ax, ay = np.abs(x), np.abs(y)
    if ax >= ay:
        c = np.float32(ay / np.float32(ax + _F_EPS))
    else:
        c = np.float32(ax / np.float32(ay + _F_EPS))
    c2 = np.float32(c * c)
    if fused:
        a = _fma(_ATAN_P7, c2, _ATAN_P5)
        a = _fma(a, c2, _ATAN_P3)
        a = _fma(a, c2, _ATAN_P1)
    else:
        a = np.float32(np.float32(_ATAN_P7 * c2) + _ATAN_P5)
        a = np.float32(np.float32(a * c2) + _ATAN_P3)
        a = np.float32(np.float32(a * c2) + _ATAN_P1)
    a = np.float32(a * c)
    if not ax >= ay:
        a = np.float32(np.float32(90.0) - a)
    if x < 0:
        a = np.float32(np.float32(180.0) - a)
    if y < 0:
        a = np.float32(np.float32(360.0) - a)
    return a


def ic_angle(ext: np.ndarray, border: int, x: int, y: int, umax: list[int], fused: bool = False) -> np.float32:
    """ICAngles: intensity-centroid orientation of the circular patch around level pixel (x, y), integer moments."""
    cy, cx = y + border, x + border
    m01 = m10 = 0
    row = ext[cy].astype(np.int64)
    for u in range(-HALF_PATCH, HALF_PATCH + 1):
        m10 += u * int(row[cx + u])
    for v in range(1, HALF_PATCH + 1):
        d = umax[v]
        us = np.arange(-d, d + 1)
        plus = ext[cy + v, cx - d:cx + d + 1].astype(np.int64)
        minus = ext[cy - v, cx - d:cx + d + 1].astype(np.int64)
        m10 += int((us * (plus + minus)).sum())
        m01 += v * int((plus - minus).sum())
    return fast_atan2(np.float32(m01), np.float32(m10), fused)


# ---- KeyPointsFilter::retainBest: libstdc++'s std::nth_element + std::partition on the responses -------------------
def _nth_element(resp: list, idx: list, first: int, nth: int, last: int) -> None:
    """std::nth_element (libstdc++ __introselect) with comparator a.response > b.response, permuting idx in place.
    resp is indexed through idx; the heap-select fallback (depth limit) is restated too."""
    def comp(i, j):          # iterators -> bool
        return resp[idx[i]] > resp[idx[j]]

    def swap(i, j):
        idx[i], idx[j] = idx[j], idx[i]

    n = last - first
    depth = 2 * (n.bit_length() - 1) if n > 0 else 0
    while last - first > 3:
        if depth == 0:
            _heap_select(resp, idx, first, nth + 1, last)
            swap(first, nth)
            return
        depth -= 1
        mid = first + (last - first) // 2
        a, b, c = first + 1, mid, last - 1
        if comp(a, b):
            if comp(b, c):
                swap(first, b)
            elif comp(a, c):
                swap(first, c)
            else:
                swap(first, a)
        elif comp(a, c):
            swap(first, a)
        elif comp(b, c):
            swap(first, c)
        else:
            swap(first, b)
        lo, hi, piv = first + 1, last, first
        while True:
            while comp(lo, piv):
                lo += 1
            hi -= 1
            while comp(piv, hi):
                hi -= 1
            if not lo < hi:
                break
            swap(lo, hi)
            lo += 1
        cut = lo
        if cut <= nth:
            first = cut
        else:
            last = cut
    # __insertion_sort(first, last)
    for i in range(first + 1, last):
        v = idx[i]
        if resp[v] > resp[idx[first]]:
            idx[first + 1:i + 1] = idx[first:i]
            idx[first] = v
        else:
            j = i
            while resp[v] > resp[idx[j - 1]]:
                idx[j] = idx[j - 1]
                j -= 1
            idx[j] = v


def _heap_select(resp, idx, first, middle, last):
    """libstdc++ __heap_select with comp = response greater (a min-heap on the response at [first, middle))."""
    def comp_v(a, b):       # on values (keypoint indices)
        return resp[a] > resp[b]

    def adjust(hole, length, value):
        top = hole
        child = hole
        while child < (length - 1) // 2:
            child = 2 * (child + 1)
            if comp_v(idx[first + child], idx[first + child - 1]):
                child -= 1
            idx[first + hole] = idx[first + child]
            hole = child
        if (length & 1) == 0 and child == (length - 2) // 2:
            child = 2 * (child + 1)
            idx[first + hole] = idx[first + child - 1]
            hole = child - 1
        parent = (hole - 1) // 2
        while hole > top and comp_v(idx[first + parent], value):
            idx[first + hole] = idx[first + parent]
            hole = parent
            parent = (hole - 1) // 2
        idx[first + hole] = value

    length = middle - first
    if length >= 2:
        parent = (length - 2) // 2
        while True:
            adjust(parent, length, idx[first + parent])
            if parent == 0:
                break
            parent -= 1
    for i in range(middle, last):
        if comp_v(idx[i], idx[first]):
            v = idx[i]
            idx[i] = idx[first]
            adjust(0, length, v)


def retain_best(resp: np.ndarray, n_points: int) -> np.ndarray:
    """KeyPointsFilter::retainBest(keypoints, n_points): the indices kept, in OpenCV's (libstdc++'s) order -- every
    keypoint whose response is >= the n_points-th largest one survives (ties included)."""
    n = len(resp)
    idx = list(range(n))
    if n_points < 0 or n <= n_points:
        return np.array(idx, dtype=np.int64)
    if n_points == 0:
        return np.zeros(0, np.int64)
    r = [float(v) for v in resp]
    _nth_element(r, idx, 0, n_points - 1, n)
    amb = r[idx[n_points - 1]]
    first, last = n_points, n          # std::partition (bidirectional form) with pred response >= amb
    while True:
        while first != last and r[idx[first]] >= amb:
            first += 1
        if first == last:
            break
        last -= 1
        while first != last and not r[idx[last]] >= amb:
            last -= 1
        if first == last:
            break
        idx[first], idx[last] = idx[last], idx[first]
        first += 1
    return np.array(idx[:first], dtype=np.int64)


def compute_keypoints(img: np.ndarray, nfeatures: int = 10000, scale_factor: float = 1.2, nlevels: int = 8,
                      edge_threshold: int = 15, fast_threshold: int = 20, fused_atan: bool = False):
    """ORB_Impl::detectAndCompute's keypoint half with FAST_SCORE: per level FAST(fast_threshold, nonmax), border filter,
    retainBest, then the orientation and the scaling to image coordinates.
    Returns (kps (n, 7) float32: x, y, size, angle, response, octave, 0;  scales, exts, border)."""
    scales, exts, border = build_pyramid(img, nlevels, scale_factor, edge_threshold)
    per_level = features_per_level(nfeatures, nlevels, scale_factor)
    umax = umax_table()
    rows_out = []
    for lv in range(nlevels):
        ext = exts[lv]
        lvl = ext[border:ext.shape[0] - border, border:ext.shape[1] - border]
        pts, resp = OF.fast_detect(lvl, fast_threshold, True)
        r, c = lvl.shape
        # runByImageBorder: keep points inside Rect(edge, edge, cols - 2 edge, rows - 2 edge) (Rect::contains: < on the far side)
        if len(pts):
            keep = ((pts[:, 0] >= edge_threshold) & (pts[:, 0] < c - edge_threshold) &
                    (pts[:, 1] >= edge_threshold) & (pts[:, 1] < r - edge_threshold))
            if r <= 2 * edge_threshold or c <= 2 * edge_threshold:
                keep[:] = False
            pts, resp = pts[keep], resp[keep]
        sel = retain_best(resp, per_level[lv])
        pts, resp = pts[sel], resp[sel]
        size = np.float32(np.float32(PATCH_SIZE) * scales[lv])
        for (x, y), rs in zip(pts, resp):
            ang = ic_angle(ext, border, int(x), int(y), umax, fused_atan)
            rows_out.append((np.float32(x) * scales[lv], np.float32(y) * scales[lv], size, ang, rs, lv, 0))
    kps = np.array(rows_out, dtype=np.float32).reshape(-1, 7)
    return kps, scales, exts, border


def compute_descriptors(kps: np.ndarray, scales: np.ndarray, exts: list, border: int, pattern: np.ndarray) -> np.ndarray:
    """computeOrbDescriptors (WTA_K 2) on the blurred pyramid: each of the 256 tests compares two steered, rounded
    sample positions around the keypoint's level pixel; bit k of byte i is test 8 i + k."""
    blurred = [gaussian_blur_7x7(e, border) for e in exts]
    pat = np.asarray(pattern, dtype=np.int64).reshape(256, 4)
    px = pat[:, [0, 2]].astype(np.float32)              # (256, 2): x of the first / second sample
    py = pat[:, [1, 3]].astype(np.float32)
    desc = np.zeros((len(kps), 32), np.uint8)
    if not len(kps):
        return desc
    lv = kps[:, 5].astype(np.int64)
    inv = (np.float32(1.0) / scales)[lv]                                   # float scale = 1.f/layerScale[octave]
    ang = (kps[:, 3] * np.float32(math.pi / 180.0)).astype(np.float32)     # angle *= (float)(CV_PI/180.f)
    a = np.cos(ang.astype(np.float64)).astype(np.float32)[:, None, None]   # (float)cos(angle)
    b = np.sin(ang.astype(np.float64)).astype(np.float32)[:, None, None]
    cy = np.rint(kps[:, 1] * inv).astype(np.int64) + border
    cx = np.rint(kps[:, 0] * inv).astype(np.int64) + border
    ix = np.rint((px[None] * a).astype(np.float32) - (py[None] * b).astype(np.float32)).astype(np.int64)
    iy = np.rint((px[None] * b).astype(np.float32) + (py[None] * a).astype(np.float32)).astype(np.int64)
    for l in range(len(exts)):
        m = lv == l
        if not m.any():
            continue
        v = blurred[l][cy[m, None, None] + iy[m], cx[m, None, None] + ix[m]]
        bits = (v[..., 0] < v[..., 1]).astype(np.uint8).reshape(-1, 32, 8)[:, :, ::-1]
        desc[m] = np.packbits(bits, axis=2).reshape(-1, 32)
    return desc


def detect_and_compute(img: np.ndarray, pattern: np.ndarray, nfeatures: int = 10000, scale_factor: float = 1.2,
                       nlevels: int = 8, edge_threshold: int = 15, fast_threshold: int = 20, fused_atan: bool = False):
    kps, scales, exts, border = compute_keypoints(img, nfeatures, scale_factor, nlevels, edge_threshold, fast_threshold,
                                                  fused_atan)
    return kps, compute_descriptors(kps, scales, exts, border, pattern)
