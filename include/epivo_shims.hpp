// epivo_shims.hpp -- header-only C++ shims with the reference drivers' call shapes over the C ABI.
//
// The reference (Ronnypetson/epivo) has no plugin interface: its drivers call
//     cv::BFMatcher(NORM_HAMMING2, true).match(desc0, desc1, matches)        kitti_ba.cpp:602,641
//     cv::findEssentialMat(p0, p1, cam, method, prob, thr, mask)             kitti_E.cpp:98-104
//     cv::recoverPose(E, p0, p1, cam, R, t, mask)                            kitti_E.cpp:120
//     Levenberg_Marquardt(n_zeta, eps, reps, wreps, lambda0, T0s, pr, p_r, lm_res)
//                                                                            jac_Rt_gen_.cpp:287-296
// on cv::Mat / std::vector<cv::Point2f> / Eigen::MatrixXd.  Neither OpenCV's nor Eigen's headers
// exist in this image, so the shims are templates over "anything that looks like" those types:
//   * a matrix type M with M(rows, cols), .rows(), .cols() and (i, j) element access
//     (Eigen::MatrixXd qualifies; tests/cpp/shim_test.cpp uses a 20-line stand-in),
//   * a 2-D point type with public .x and .y floats (cv::Point2f qualifies).
// A driver switches by including this header, linking libepivo_b200.so and replacing the
// `cv::` / unqualified calls with `epivo::` ones -- see INTEGRATION.md.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <stdexcept>
#include <algorithm>
#include <cmath>
#include <string>
#include <utility>
#include <vector>

#include "epivo_b200.h"

// The reference uses `LM_res` (jac_Rt_gen_.cpp:295,473-475; kitti_ba.cpp:876) but defines it
// nowhere; this is the definition its fields imply.
struct LM_res {
    double H_norm = 0.0, r_norm = 0.0, lambda = 0.0;
};

namespace epivo {

enum { NORM_HAMMING = EPIVO_NORM_HAMMING, NORM_HAMMING2 = EPIVO_NORM_HAMMING2, LMEDS = EPIVO_LMEDS, RANSAC = EPIVO_RANSAC };

// One context per host thread (kitti_ba.cpp:1153-1163 runs association and BA on different
// threads concurrently: give each its own Context).
class Context {
  public:
    explicit Context(int device = 0) {
        if (epivo_create(&ctx_, device) != EPIVO_OK || !ctx_)
            throw std::runtime_error("epivo_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { epivo_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    epivo_ctx* get() const { return ctx_; }
    void check(int rc) const {
        if (rc != EPIVO_OK) throw std::runtime_error(std::string("epivo_b200: ") + epivo_last_error(ctx_));
    }

  private:
    epivo_ctx* ctx_ = nullptr;
};

// One context per host thread, created on first use (device from $EPIVO_DEVICE, default 0): what the
// handle-free overloads of include/epivo_dropin.hpp run on.  Throws if there is no CUDA device.
inline Context& default_context() {
    thread_local Context ctx([] {
        const char* e = std::getenv("EPIVO_DEVICE");
        return e ? std::atoi(e) : 0;
    }());
    return ctx;
}

// How to read / create a matrix of the caller's type.  Default: Eigen-like (rows(), cols(), (i, j), M(r, c)).
template <typename MatT, typename Enable = void>
struct mat_traits {
    static int rows(const MatT& m) { return (int)m.rows(); }
    static int cols(const MatT& m) { return (int)m.cols(); }
    static bool empty(const MatT& m) { return m.rows() == 0 || m.cols() == 0; }
    static double get(const MatT& m, int i, int j) { return (double)m(i, j); }
    static MatT create(int r, int c) { return MatT(r, c); }          // double entries
    static void set(MatT& m, int i, int j, double v) { m(i, j) = v; }
    static const unsigned char* bytes(const MatT& m) { return reinterpret_cast<const unsigned char*>(m.data()); }
};

#ifdef OPENCV_CORE_MAT_HPP
// cv::Mat: element type by depth (the reference's `cam` is CV_32F, kitti_E.cpp:38; E / R / t come back CV_64F,
// as from OpenCV); public rows / cols / data members.
template <>
struct mat_traits<cv::Mat, void> {
    static int rows(const cv::Mat& m) { return m.rows; }
    static int cols(const cv::Mat& m) { return m.cols; }
    static bool empty(const cv::Mat& m) { return m.empty(); }
    static double get(const cv::Mat& m, int i, int j) {
        return m.depth() == CV_32F ? (double)m.at<float>(i, j) : m.at<double>(i, j);
    }
    static cv::Mat create(int r, int c) { return cv::Mat(r, c, CV_64F); }
    static void set(cv::Mat& m, int i, int j, double v) { m.at<double>(i, j) = v; }
    static const unsigned char* bytes(const cv::Mat& m) { return m.data; }
    static cv::Mat create_u8(int r, int c) { return cv::Mat(r, c, CV_8U); }      // the output image of remap
    static unsigned char* bytes_mut(cv::Mat& m) { return m.data; }
};
#endif

struct DMatch {               // cv::DMatch
    int queryIdx = -1, trainIdx = -1, imgIdx = 0;
    float distance = 0.f;
};

// cv::BFMatcher(normType, crossCheck); descriptors as a contiguous row-major byte matrix
class BFMatcher {
  public:
    BFMatcher(Context& ctx, int normType = NORM_HAMMING2, bool crossCheck = false)
        : ctx_(ctx), norm_(normType), cross_(crossCheck) {}
    // cv::BFMatcher's own constructor (kitti_ba.cpp:602): runs on the calling thread's default context
    explicit BFMatcher(int normType = NORM_HAMMING2, bool crossCheck = false)
        : ctx_(default_context()), norm_(normType), cross_(crossCheck) {}
    // desc0 / desc1: n x desc_bytes, row-major uint8 (cv::Mat::data of an ORB descriptor matrix)
    void match(const uint8_t* desc0, int n0, const uint8_t* desc1, int n1, int desc_bytes,
               std::vector<DMatch>& matches) const {
        std::vector<int32_t> q(n0 > 0 ? n0 : 1), t(q.size()), d(q.size());
        int n = 0;
        ctx_.check(epivo_match_hamming(ctx_.get(), desc0, n0, desc1, n1, desc_bytes, norm_,
                                       cross_ ? EPIVO_MATCH_CROSSCHECK : EPIVO_MATCH_NN, 0.f, q.data(), t.data(),
                                       d.data(), nullptr, &n));
        matches.resize(n);
        for (int i = 0; i < n; ++i) {
            matches[i].queryIdx = q[i];
            matches[i].trainIdx = t[i];
            matches[i].imgIdx = 0;
            matches[i].distance = (float)d[i];
        }
    }

    // matcher.match(desc0, desc1, matches) on the descriptor MATRICES themselves (kitti_ba.cpp:641): MatT is a cv::Mat
    // of CV_8U rows (or anything mat_traits can read bytes from), DM anything with cv::DMatch's four fields.
    template <typename MatT, typename DM>
    void match(const MatT& queryDescriptors, const MatT& trainDescriptors, std::vector<DM>& matches) const {
        typedef mat_traits<MatT> MT;
        if (!MT::empty(queryDescriptors) && !MT::empty(trainDescriptors) &&
            MT::cols(queryDescriptors) != MT::cols(trainDescriptors))
            throw std::invalid_argument("descriptor sizes differ");
        std::vector<DMatch> m;
        match(MT::bytes(queryDescriptors), MT::rows(queryDescriptors), MT::bytes(trainDescriptors),
              MT::rows(trainDescriptors), MT::cols(queryDescriptors), m);
        matches.resize(m.size());
        for (size_t i = 0; i < m.size(); ++i) {
            matches[i].queryIdx = m[i].queryIdx;
            matches[i].trainIdx = m[i].trainIdx;
            matches[i].imgIdx = m[i].imgIdx;
            matches[i].distance = m[i].distance;
        }
    }

  private:
    Context& ctx_;
    int norm_;
    bool cross_;
};

namespace detail {
template <typename Pt>
std::vector<float> flatten(const std::vector<Pt>& p) {
    std::vector<float> o(2 * p.size());
    for (size_t i = 0; i < p.size(); ++i) { o[2 * i] = p[i].x; o[2 * i + 1] = p[i].y; }
    return o;
}
}  // namespace detail

// cv::findEssentialMat(points1, points2, cameraMatrix, method, prob, threshold, mask).
// cam: 9 doubles, row-major (widen the reference's float Mat: kitti_E.cpp:38).  Returns false where
// OpenCV returns an empty Mat (fewer than 5 points / no model); E is 3x3 row-major.
template <typename Pt>
bool findEssentialMat(Context& ctx, const std::vector<Pt>& points1, const std::vector<Pt>& points2,
                      const double cam[9], int method, double prob, double threshold, double E[9],
                      std::vector<unsigned char>& mask, int maxIters = 1000) {
    if (points1.size() != points2.size()) throw std::invalid_argument("point sets differ in size");
    const int n = (int)points1.size();
    std::vector<float> p0 = detail::flatten(points1), p1 = detail::flatten(points2);
    mask.assign(n, 0);
    int ninl = 0, iters = 0;
    int rc = epivo_find_essential(ctx.get(), p0.data(), p1.data(), n, cam, method, prob, threshold, maxIters, nullptr, 0,
                                  E, mask.data(), &ninl, &iters);
    if (rc == EPIVO_ERR_NOMODEL) return false;
    ctx.check(rc);
    return true;
}

// cv::recoverPose(E, points1, points2, cameraMatrix, R, t, mask): returns the number of points that
// pass the cheirality check; mask entries are 0 / 255 (the drivers test == 255, kitti_E.cpp:177).
template <typename Pt>
int recoverPose(Context& ctx, const double E[9], const std::vector<Pt>& points1, const std::vector<Pt>& points2,
                const double cam[9], double R[9], double t[3], std::vector<unsigned char>& mask,
                double distanceThresh = 50.0) {
    const int n = (int)points1.size();
    std::vector<float> p0 = detail::flatten(points1), p1 = detail::flatten(points2);
    mask.assign(n, 0);
    int good = 0;
    ctx.check(epivo_recover_pose(ctx.get(), E, p0.data(), p1.data(), n, cam, distanceThresh, nullptr, R, t,
                                 mask.data(), &good));
    return good;
}

// int Levenberg_Marquardt(const int n_zeta, const double epsilon, const vector<pair<int,int>>& reps,
//                         const vector<double>& wreps, const double lambda0, vector<MatrixXd>& T0s,
//                         vector<MatrixXd>& pr, vector<MatrixXd>& p_r, LM_res& lm_res)
// (jac_Rt_gen_.cpp:287-296).  T0s (4x4 each) is updated in place, as in the reference.
template <typename M>
int Levenberg_Marquardt(Context& ctx, const int n_zeta, const double epsilon,
                        const std::vector<std::pair<int, int> >& reps, const std::vector<double>& wreps,
                        const double lambda0, std::vector<M>& T0s, std::vector<M>& pr, std::vector<M>& p_r,
                        LM_res& lm_res, double huber_delta = 1e-5 /* jac_Rt_gen_.cpp:17 */, int max_iters = 30) {
    if (reps.size() != wreps.size()) throw std::invalid_argument("reps.size() != wreps.size()");   // :297
    const int n_rep = (int)reps.size();
    if ((int)T0s.size() != n_zeta || (int)pr.size() != n_rep || (int)p_r.size() != n_rep)
        throw std::invalid_argument("T0s / pr / p_r sizes do not match n_zeta / reps");
    const int N = (int)pr[0].rows();                                                               // :299
    std::vector<int32_t> r(2 * n_rep);
    for (int j = 0; j < n_rep; ++j) { r[2 * j] = reps[j].first; r[2 * j + 1] = reps[j].second; }
    std::vector<double> T(16 * (size_t)n_zeta), a(3 * (size_t)n_rep * N), b(a.size());
    for (int k = 0; k < n_zeta; ++k)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) T[16 * k + 4 * i + j] = T0s[k](i, j);
    for (int j = 0; j < n_rep; ++j) {
        if ((int)pr[j].rows() != N || (int)p_r[j].rows() != N) throw std::invalid_argument("all reps share one N");
        for (int i = 0; i < N; ++i)
            for (int c = 0; c < 3; ++c) {
                a[((size_t)j * N + i) * 3 + c] = pr[j](i, c);
                b[((size_t)j * N + i) * 3 + c] = p_r[j](i, c);
            }
    }
    epivo_lm_res res;
    int iters = 0;
    ctx.check(epivo_lm_rt(ctx.get(), n_zeta, epsilon, r.data(), wreps.data(), n_rep, lambda0, max_iters, huber_delta,
                          T.data(), a.data(), b.data(), N, &res, &iters));
    for (int k = 0; k < n_zeta; ++k)
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) T0s[k](i, j) = T[16 * k + 4 * i + j];
    lm_res.H_norm = res.H_norm;
    lm_res.r_norm = res.r_norm;
    lm_res.lambda = res.lambda;
    return 0;
}

// The 7-argument form kitti_E.cpp:196 calls -- unit weights, one diagnostic returned.  The reference
// never defines this overload; of the three candidates it leaves commented out
// (jac_Rt_gen_.cpp:470-472: H.norm(), r0.norm(), lambda) the residual norm is returned, which is what
// the caller's `uncert > 1E-9` revert test (kitti_E.cpp:198) is dimensionally consistent with.
template <typename M>
double Levenberg_Marquardt(Context& ctx, const int n_zeta, const double epsilon,
                           const std::vector<std::pair<int, int> >& reps, const double lambda0, std::vector<M>& T0s,
                           std::vector<M>& pr, std::vector<M>& p_r) {
    std::vector<double> w(reps.size(), 1.0);
    LM_res res;
    Levenberg_Marquardt(ctx, n_zeta, epsilon, reps, w, lambda0, T0s, pr, p_r, res);
    return res.r_norm;
}

// `struct reproj` (kitti_ba.cpp:167-175) with plain arrays for R (row-major) and t.
template <typename Pt>
struct Reproj {
    std::vector<Pt> p0, p1;
    double R[9], t[3];
    double w;                 // reprojection weight (kitti_ba.cpp:171-172): 0 freezes it (the stereo extrinsics)
    Reproj() : w(1.0) {
        for (int i = 0; i < 9; ++i) R[i] = i % 4 == 0 ? 1.0 : 0.0;
        t[0] = t[1] = t[2] = 0.0;
    }
};

// void match_kp(window, stride, num_frames, source_kp, img_fns, descs, cam, reprojs)   kitti_ba.cpp:583-755
// for a whole drive in ONE device pass.  The reference walks (i + window[j].first, i + window[j].second) pair by
// pair on the association thread (BFMatcher NORM_HAMMING2 cross-check :602,641; findEssentialMat(LMEDS, 0.99, 0.1)
// :702; mask == 1 compaction :705-710; recoverPose :715; rec_mask == 255 compaction :729-735; fewer than 8
// matches -> identity and (0.1, 0.1, -0.9), :741-744).  Here the walk becomes an explicit pair list
// (epivo_seq_set_pairs) over the frames' keypoints / descriptors and the map is filled from the results.
//   source_kp[f]  keypoints of frame f (any count per frame)
//   descs[f]      source_kp[f].size() x 32 descriptor bytes, row-major (cv::Mat::data of the ORB descriptors)
template <typename Pt>
void match_kp(Context& ctx, const std::vector<std::pair<int, int> >& window, const int stride, const int num_frames,
              const std::vector<std::vector<Pt> >& source_kp, const std::vector<const uint8_t*>& descs,
              const double cam[9], std::map<std::pair<int, int>, Reproj<Pt> >& reprojs) {
    if (stride <= 0) throw std::invalid_argument("stride must be positive");                       // :592 assert
    if ((int)source_kp.size() < num_frames || (int)descs.size() < num_frames)
        throw std::invalid_argument("source_kp / descs shorter than num_frames");
    std::vector<std::pair<int, int> > pairs;                                                       // :603-615
    for (int i = 0; i < num_frames; i += stride) {
        for (size_t j = 0; j < window.size(); ++j) {
            const std::pair<int, int> key(i + window[j].first, i + window[j].second);
            if (reprojs.count(key)) continue;
            bool seen = false;
            for (size_t q = 0; q < pairs.size() && !seen; ++q) seen = pairs[q] == key;
            if (seen) continue;
            if (key.first >= num_frames || key.second >= num_frames) break;
            pairs.push_back(key);
        }
    }
    if (pairs.empty() || num_frames < 2) return;
    size_t kp = 1;
    for (int f = 0; f < num_frames; ++f) kp = source_kp[f].size() > kp ? source_kp[f].size() : kp;
    std::vector<float> kps((size_t)num_frames * kp * 2, 0.f);
    std::vector<uint8_t> dsc((size_t)num_frames * kp * 32, 0);
    std::vector<int32_t> counts(num_frames), fq(pairs.size()), ft(pairs.size());
    for (int f = 0; f < num_frames; ++f) {
        counts[f] = (int32_t)source_kp[f].size();
        for (size_t k = 0; k < source_kp[f].size(); ++k) {
            kps[((size_t)f * kp + k) * 2] = source_kp[f][k].x;
            kps[((size_t)f * kp + k) * 2 + 1] = source_kp[f][k].y;
        }
        if (counts[f]) std::copy(descs[f], descs[f] + (size_t)counts[f] * 32, dsc.begin() + (size_t)f * kp * 32);
    }
    for (size_t p = 0; p < pairs.size(); ++p) { fq[p] = pairs[p].first; ft[p] = pairs[p].second; }
    epivo_seq* seq = nullptr;
    const int max_pairs = (int)pairs.size() > num_frames - 1 ? (int)pairs.size() : num_frames - 1;
    ctx.check(epivo_seq_create_pairs(ctx.get(), &seq, num_frames, (int)kp, max_pairs));
    struct Guard { epivo_seq* s; ~Guard() { epivo_seq_destroy(s); } } guard = {seq};
    ctx.check(epivo_seq_set_counts(seq, 0, num_frames, counts.data()));
    ctx.check(epivo_seq_set_pairs(seq, (int)pairs.size(), fq.data(), ft.data()));
    epivo_pipeline_params prm;
    epivo_pipeline_params_default(&prm);
    for (int i = 0; i < 9; ++i) prm.K[i] = cam[i];
    prm.method = EPIVO_LMEDS;                                                                      // :702
    prm.prob = 0.99;
    prm.threshold = 0.1;
    std::vector<epivo_pair_result> res(pairs.size());
    ctx.check(epivo_seq_process(seq, &prm, num_frames, kps.data(), dsc.data(), res.data()));
    std::vector<int32_t> qi(kp), ti(kp), di(kp);
    std::vector<uint8_t> em(kp), pm(kp);
    for (size_t p = 0; p < pairs.size(); ++p) {
        Reproj<Pt> rep;
        const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, t0[3] = {0.1, 0.1, -0.9};                 // :741-744
        std::copy(I, I + 9, rep.R);
        std::copy(t0, t0 + 3, rep.t);
        // (no model -- findEssentialMat would have returned an empty Mat, on which the reference's recoverPose
        // throws -- keeps the identity / (0.1, 0.1, -0.9) start instead of the zeroed pose-stage outputs)
        if (res[p].n_matches >= 8 && res[p].n_inliers > 0) {                                       // :700
            int nm = 0, ne = 0, np = 0;
            ctx.check(epivo_seq_get_matches(seq, (int)p, qi.data(), ti.data(), di.data(), &nm));
            ctx.check(epivo_seq_get_masks(seq, (int)p, em.data(), &ne, pm.data(), &np));
            const std::vector<Pt>& k0 = source_kp[pairs[p].first];
            const std::vector<Pt>& k1 = source_kp[pairs[p].second];
            int c = 0;                                             // index into the compacted E-inlier list
            for (int m = 0; m < nm; ++m) {
                if (em[m] != 1) continue;                                                          // :705-710
                if (c < np && pm[c] == 255) {                                                      // :729-735
                    rep.p0.push_back(k0[qi[m]]);
                    rep.p1.push_back(k1[ti[m]]);
                }
                ++c;
            }
            std::copy(res[p].R, res[p].R + 9, rep.R);
            std::copy(res[p].t, res[p].t + 3, rep.t);
        }
        reprojs.insert(std::make_pair(pairs[p], rep));
    }
}

// int bundle_adjustment(reprojs, window, stride, num_frames, cam, opt_T)                    kitti_ba.cpp:757-905
// The reference walks the windows one by one, polling the reprojs map with 20 ms sleeps (:793-797).  A window's
// LM problem depends only on that window's reprojections -- its initial chain is rebuilt from reprojs[(j, j+1)]
// every time (:856-859, `!optimized[j] || true`) -- so all windows are assembled first and solved by ONE
// epivo_lm_rt_batch launch; only the sequential tail is replayed on the host as written: revert when
// r_norm > 1e-2 (:889-891), divide the translations by the scale carried from the previous window (:853-855,
// :898-901), later windows overwrite the overlap.  Every reprojection a window needs must already be in the map
// (std::out_of_range otherwise: there is no association thread to wait for -- see match_kp above).
//   M: a matrix type with M(rows, cols) and (i, j) access (Eigen::MatrixXd); opt_T must come in empty (:764).
namespace detail {
// mono: nodes = frames, node_step 1, unit weights, scale carry; stereo: nodes 2f (left) / 2f + 1 (right), node_step 2,
// the reprojections' own weights, no scale carry (kitti_ba.cpp:908-1068)
template <typename Pt, typename M>
int windowed_ba(Context& ctx, const std::map<std::pair<int, int>, Reproj<Pt> >& reprojs,
                const std::vector<std::pair<int, int> >& window, const int stride, const int num_frames,
                const int node_step, const double cam[9], std::vector<M>& opt_T, std::vector<LM_res>* lm_out,
                double huber_delta) {
    const bool stereo = node_step == 2;
    const int num_nodes = node_step * num_frames;
    if (stride <= 0) throw std::invalid_argument("stride must be positive");                       // :763
    if (!opt_T.empty()) throw std::invalid_argument("opt_T must be empty");                        // :764
    if (window.empty()) throw std::invalid_argument("empty window");
    for (int f = 0; f < num_nodes; ++f) {                                                          // :767-770, :918-921
        M I(4, 4);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) I(i, j) = i == j ? 1.0 : 0.0;
        opt_T.push_back(I);
    }
    if (lm_out) lm_out->clear();
    // cam^-1 (:772-774), general 3x3 inverse
    double Ki[9];
    {
        const double* k = cam;
        const double det = k[0] * (k[4] * k[8] - k[5] * k[7]) - k[1] * (k[3] * k[8] - k[5] * k[6]) + k[2] * (k[3] * k[7] - k[4] * k[6]);
        Ki[0] = (k[4] * k[8] - k[5] * k[7]) / det; Ki[1] = (k[2] * k[7] - k[1] * k[8]) / det; Ki[2] = (k[1] * k[5] - k[2] * k[4]) / det;
        Ki[3] = (k[5] * k[6] - k[3] * k[8]) / det; Ki[4] = (k[0] * k[8] - k[2] * k[6]) / det; Ki[5] = (k[2] * k[3] - k[0] * k[5]) / det;
        Ki[6] = (k[3] * k[7] - k[4] * k[6]) / det; Ki[7] = (k[1] * k[6] - k[0] * k[7]) / det; Ki[8] = (k[0] * k[4] - k[1] * k[3]) / det;
    }
    const int min_pt = 32, n_rep = (int)window.size();                                             // :777
    int lo = window[0].first, hi = window[0].first;
    std::vector<int32_t> reps(2 * (size_t)n_rep);
    for (int j = 0; j < n_rep; ++j) {
        const int a = window[j].first, b = window[j].second;
        if (a == b) throw std::invalid_argument("window entry with first == second");              // :812 assert
        lo = std::min(lo, std::min(a, b));
        hi = std::max(hi, std::max(a, b));
        reps[2 * j] = b > a ? a : a - 1;                                                           // :813-817
        reps[2 * j + 1] = b > a ? b - 1 : b;
    }
    const int nz = hi - lo;                                                                        // w1 - w0 (:874)
    std::vector<int> starts;                                                                       // :780-801
    for (int i = 0; i < num_frames; i += stride) {
        if (node_step * i + hi >= num_nodes) break;
        starts.push_back(i);
    }
    const int B = (int)starts.size();
    if (B == 0) return 0;
    std::vector<double> T0((size_t)B * nz * 16, 0.0), pr((size_t)B * n_rep * min_pt * 3, 1.0), p_r(pr.size(), 1.0),
        w((size_t)B * n_rep, 0.0);
    for (int b = 0; b < B; ++b) {
        const int i = node_step * starts[b];
        for (int j = 0; j < n_rep; ++j) {
            const Reproj<Pt>& r = reprojs.at(std::make_pair(i + window[j].first, i + window[j].second));
            if ((int)r.p0.size() < min_pt) continue;                // "Bad pts": weight 0, all-ones points (:819-824)
            w[(size_t)b * n_rep + j] = stereo ? r.w : 1.0;                                         // :832 / :996
            for (int k = 0; k < min_pt; ++k) {                                                     // :838-845
                const double u0 = r.p0[k].x, v0 = r.p0[k].y, u1 = r.p1[k].x, v1 = r.p1[k].y;
                double* a = &pr[(((size_t)b * n_rep + j) * min_pt + k) * 3];
                double* c = &p_r[(((size_t)b * n_rep + j) * min_pt + k) * 3];
                for (int q = 0; q < 3; ++q) {
                    a[q] = Ki[3 * q] * u0 + Ki[3 * q + 1] * v0 + Ki[3 * q + 2];
                    c[q] = Ki[3 * q] * u1 + Ki[3 * q + 1] * v1 + Ki[3 * q + 2];
                }
            }
        }
        for (int k = 0; k < nz; ++k) {                                                             // :856-868
            const Reproj<Pt>& r = reprojs.at(std::make_pair(i + lo + k, i + lo + k + 1));
            double* T = &T0[((size_t)b * nz + k) * 16];
            for (int q = 0; q < 3; ++q) {
                for (int c = 0; c < 3; ++c) T[4 * q + c] = r.R[3 * q + c];
                T[4 * q + 3] = r.t[q];
            }
            T[15] = 1.0;
        }
    }
    std::vector<double> T(T0);
    std::vector<epivo_lm_res> res(B);
    std::vector<int32_t> iters(B);
    ctx.check(epivo_lm_rt_batch(ctx.get(), B, nz, 1e-8, reps.data(), w.data(), n_rep, 1e-2, 30, huber_delta, T.data(),
                                pr.data(), p_r.data(), min_pt, res.data(), iters.data()));        // :881
    std::vector<bool> optimized(num_nodes, false);
    for (int b = 0; b < B; ++b) {                                                                  // :853-855, :889-903
        const int w0 = node_step * starts[b] + lo;
        double scale = 1.0;
        if (!stereo && optimized[w0]) {
            const M& P = opt_T[w0];
            scale = std::sqrt(P(0, 3) * P(0, 3) + P(1, 3) * P(1, 3) + P(2, 3) * P(2, 3));
        }
        const double* Ts = res[b].r_norm > 1e-2 ? &T0[(size_t)b * nz * 16] : &T[(size_t)b * nz * 16];
        for (int k = 0; k < nz; ++k) {
            M& O = opt_T[w0 + k];
            for (int q = 0; q < 4; ++q) for (int c = 0; c < 4; ++c) O(q, c) = Ts[16 * k + 4 * q + c];
            for (int q = 0; q < 3; ++q) O(q, 3) /= scale;
            optimized[w0 + k] = true;
        }
        if (lm_out) {
            LM_res l;
            l.H_norm = res[b].H_norm; l.r_norm = res[b].r_norm; l.lambda = res[b].lambda;
            lm_out->push_back(l);
        }
    }
    return B;
}
}  // namespace detail

template <typename Pt, typename M>
int bundle_adjustment(Context& ctx, const std::map<std::pair<int, int>, Reproj<Pt> >& reprojs,
                      const std::vector<std::pair<int, int> >& window, const int stride, const int num_frames,
                      const double cam[9], std::vector<M>& opt_T, std::vector<LM_res>* lm_out = nullptr,
                      double huber_delta = 1e-5 /* jac_Rt_gen_.cpp:17 */) {
    return detail::windowed_ba(ctx, reprojs, window, stride, num_frames, 1, cam, opt_T, lm_out, huber_delta);
}

// int bundle_adjustment_stereo(reprojs, window, stride, num_frames, cam, opt_T)             kitti_ba.cpp:908-1068
// -- the form the shipped main() runs (:1157).  Frame f becomes nodes 2f (left) and 2f + 1 (right); every window
// entry (a, b) expands to (2a, 2b), (2a + 1, 2b), (2a, 2a + 1) (:934-941); the reprojections carry their own
// weights (0 on left -> right: frozen extrinsics); opt_T gets 2 * num_frames entries; no scale carry.
template <typename Pt, typename M>
int bundle_adjustment_stereo(Context& ctx, const std::map<std::pair<int, int>, Reproj<Pt> >& reprojs,
                             const std::vector<std::pair<int, int> >& window, const int stride, const int num_frames,
                             const double cam[9], std::vector<M>& opt_T, std::vector<LM_res>* lm_out = nullptr,
                             double huber_delta = 1e-5) {
    std::vector<std::pair<int, int> > ws;
    for (size_t i = 0; i < window.size(); ++i) {
        const int a = window[i].first, b = window[i].second;
        ws.push_back(std::make_pair(2 * a, 2 * b));
        ws.push_back(std::make_pair(2 * a + 1, 2 * b));
        ws.push_back(std::make_pair(2 * a, 2 * a + 1));
    }
    return detail::windowed_ba(ctx, reprojs, ws, stride, num_frames, 2, cam, opt_T, lm_out, huber_delta);
}

}  // namespace epivo
