// epivo_dropin.hpp -- the reference's call lines, VERBATIM, over libepivo_b200.
//
// include/epivo_shims.hpp exposes the path with an explicit epivo::Context and raw arrays.  This header
// adds overloads with exactly the signatures the reference's drivers call, so that a driver switches by
// adding ONE include after its own OpenCV / Eigen includes and touching no call site:
//
//   Mat ess = findEssentialMat(_cpt0, _cpt1_, cam, LMEDS, 0.99, 0.01, mask_ess);        kitti_E.cpp:98-104
//   recoverPose(ess, cpt0, cpt1_, cam, rot, tr, rec_mask);                              kitti_E.cpp:120
//   uncert = Levenberg_Marquardt(1, 1e-8, reps, 1e-2, T0s, pr, p_r);                    kitti_E.cpp:196
//   Levenberg_Marquardt(nzeta, 1e-8, reps, wreps, 1e-2, T0s, pr, p_r, lm_res);          kitti_ba.cpp:881,1044
//                        -- int Levenberg_Marquardt(const int, const double, const vector<pair<int,int>>&,
//                           const vector<double>&, const double, vector<MatrixXd>&, vector<MatrixXd>&,
//                           vector<MatrixXd>&, LM_res&)                                 jac_Rt_gen_.cpp:287-296
//   BFMatcher matcher(NORM_HAMMING2, true);  matcher.match(desc0, desc1, matches);      kitti_ba.cpp:602,641
//   Ptr<FastFeatureDetector> detector = FastFeatureDetector::create(40);                kitti_E.cpp:70
//   detector->detect(src, kp0, Mat());                                                  kitti_E.cpp:73
//   calcOpticalFlowPyrLK(src, tgt, pt0, pt1_, status, err);                             kitti_E.cpp:79-84
//   remap(src, src_, map1, map2, INTER_LINEAR);                                         euroc_E.cpp:170
//   Ptr<ORB> orb = ORB::create(10000, 1.2f, 8, 15, 0, 2, ORB::FAST_SCORE);              kitti_ba.cpp:128
//   orb->detect(src, kp0, Mat());  orb->compute(src, kp0, desc0);                       kitti_ba.cpp:141,149
//
// How the unqualified calls reach these functions.  The drivers say `using namespace cv;` and call
// `findEssentialMat(...)` unqualified.  The templates below live in the GLOBAL namespace and take the
// caller's own types (std::vector<cv::Point2f>, cv::Mat, std::vector<uchar>) exactly, whereas
// cv::findEssentialMat / cv::recoverPose take InputArray / OutputArray, which need a user-defined
// conversion from every argument: overload resolution picks the exact match, i.e. the template, without
// any edit.  A class cannot be overloaded that way, so for the matcher the driver writes
// `epivo::BFMatcher` instead of `BFMatcher` (one token, kitti_ba.cpp:602), and for the detector
// `epivo::FastFeatureDetector` in the declaration and the create() call (kitti_E.cpp:70), and likewise
// `epivo::ORB` (kitti_ba.cpp:128).
//
// Context.  The calls carry no handle, so each host thread gets its own lazily created epivo::Context
// (`thread_local`; device from $EPIVO_DEVICE, default 0) -- kitti_ba.cpp:1153-1163 calls the path from
// two std::threads concurrently, and a context must not be shared between threads.
//
// Matrix types.  `MatT` is anything epivo::mat_traits knows: cv::Mat (specialisation below, compiled when
// OpenCV's core header has been included first) or a type with M(rows, cols) / rows() / cols() / (i, j)
// such as Eigen::MatrixXd.  Neither OpenCV nor Eigen exists in this image: tests/cpp/dropin_test.cpp
// compiles the literal call lines above against stand-ins with the same member API.
#pragma once
#include <cstring>
#include <memory>
#include <type_traits>

#include "epivo_shims.hpp"

namespace epivo {

namespace detail {
template <typename MatT>
void cam_to_array(const MatT& cam, double K[9]) {
    typedef mat_traits<MatT> MT;
    if (MT::rows(cam) != 3 || MT::cols(cam) != 3) throw std::invalid_argument("cameraMatrix must be 3x3");
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) K[3 * i + j] = MT::get(cam, i, j);
}
}  // namespace detail

}  // namespace epivo

// ---- the reference's unqualified calls ---------------------------------------------------------------------------

// Mat findEssentialMat(points1, points2, cameraMatrix, method, prob, threshold, mask)        kitti_E.cpp:98-104
// Returns a 3x3 CV_64F-like matrix, or an empty (0x0) one where OpenCV returns an empty Mat; mask in {0,1}.
template <typename Pt, typename MatT>
MatT findEssentialMat(const std::vector<Pt>& points1, const std::vector<Pt>& points2, const MatT& cameraMatrix,
                      int method, double prob, double threshold, std::vector<unsigned char>& mask) {
    typedef epivo::mat_traits<MatT> MT;
    double K[9], E[9];
    epivo::detail::cam_to_array(cameraMatrix, K);
    if (!epivo::findEssentialMat(epivo::default_context(), points1, points2, K, method, prob, threshold, E, mask))
        return MT::create(0, 0);
    MatT out = MT::create(3, 3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) MT::set(out, i, j, E[3 * i + j]);
    return out;
}

// int recoverPose(E, points1, points2, cameraMatrix, R, t, mask)                             kitti_E.cpp:120
// R (3x3) and t (3x1) are created; mask in {0,255} (the drivers test == 255, kitti_E.cpp:177).
template <typename Pt, typename MatT>
int recoverPose(const MatT& E, const std::vector<Pt>& points1, const std::vector<Pt>& points2, const MatT& cameraMatrix,
                MatT& R, MatT& t, std::vector<unsigned char>& mask) {
    typedef epivo::mat_traits<MatT> MT;
    if (MT::rows(E) != 3 || MT::cols(E) != 3) throw std::invalid_argument("E must be 3x3");
    double K[9], e[9], r[9], tt[3];
    epivo::detail::cam_to_array(cameraMatrix, K);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) e[3 * i + j] = MT::get(E, i, j);
    const int good = epivo::recoverPose(epivo::default_context(), e, points1, points2, K, r, tt, mask);
    R = MT::create(3, 3);
    t = MT::create(3, 1);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) MT::set(R, i, j, r[3 * i + j]);
        MT::set(t, i, 0, tt[i]);
    }
    return good;
}

// int Levenberg_Marquardt(n_zeta, epsilon, reps, wreps, lambda0, T0s, pr, p_r, lm_res)       jac_Rt_gen_.cpp:287-296
// The reference's huber_delta is a compile-time constant of that file (1e-5, :17) and so it is here.
template <typename M>
int Levenberg_Marquardt(const int n_zeta, const double epsilon, const std::vector<std::pair<int, int> >& reps,
                        const std::vector<double>& wreps, const double lambda0, std::vector<M>& T0s, std::vector<M>& pr,
                        std::vector<M>& p_r, LM_res& lm_res) {
    return epivo::Levenberg_Marquardt(epivo::default_context(), n_zeta, epsilon, reps, wreps, lambda0, T0s, pr, p_r, lm_res);
}

// double Levenberg_Marquardt(n_zeta, epsilon, reps, lambda0, T0s, pr, p_r)                   kitti_E.cpp:196
// The 7-argument call the reference makes but never defines; returns the residual norm (see epivo_shims.hpp).
template <typename M>
double Levenberg_Marquardt(const int n_zeta, const double epsilon, const std::vector<std::pair<int, int> >& reps,
                           const double lambda0, std::vector<M>& T0s, std::vector<M>& pr, std::vector<M>& p_r) {
    return epivo::Levenberg_Marquardt(epivo::default_context(), n_zeta, epsilon, reps, lambda0, T0s, pr, p_r);
}

// ---- N4 front end: detector and tracker --------------------------------------------------------------------------
namespace epivo {

// cv::FastFeatureDetector (TYPE_9_16).  `Ptr<epivo::FastFeatureDetector> detector = epivo::FastFeatureDetector::create(40);`
// then `detector->detect(src, kp0, Mat());` unchanged (kitti_E.cpp:70-73, kitti_ba.cpp:49-62,98-118).  create() returns a
// std::shared_ptr, which cv::Ptr is constructible from.  KP is the caller's cv::KeyPoint: pt, size (7), angle (-1),
// response (the corner score), octave (0), class_id (-1) are set as OpenCV sets them.
class FastFeatureDetector {
  public:
    static std::shared_ptr<FastFeatureDetector> create(int threshold = 10, bool nonmaxSuppression = true) {
        return std::shared_ptr<FastFeatureDetector>(new FastFeatureDetector(threshold, nonmaxSuppression));
    }
    // image: 8-bit single channel, rows x cols, contiguous (a freshly imread grey image is); the mask must be empty
    template <typename MatT, typename KP>
    void detect(const MatT& image, std::vector<KP>& keypoints, const MatT& mask = MatT()) const {
        typedef mat_traits<MatT> MT;
        if (!MT::empty(mask)) throw std::invalid_argument("FastFeatureDetector::detect: masks are not supported");
        keypoints.clear();
        if (MT::empty(image)) return;
        Context& ctx = default_context();
        const int rows = MT::rows(image), cols = MT::cols(image);
        int cap = std::max(1024, rows * cols / 16);
        std::vector<float> xy, resp;
        int32_t found = 0;
        for (;;) {                              // a frame with more corners than the buffer holds is run again
            xy.assign(2 * (size_t)cap, 0.f);
            resp.assign((size_t)cap, 0.f);
            ctx.check(epivo_fast_detect(ctx.get(), MT::bytes(image), 1, rows, cols, threshold_, nonmax_ ? 1 : 0, cap,
                                        xy.data(), resp.data(), &found));
            if (found <= cap) break;
            cap = found;
        }
        keypoints.resize((size_t)found);
        for (int i = 0; i < found; ++i) {
            KP& k = keypoints[(size_t)i];
            k.pt.x = xy[2 * (size_t)i];
            k.pt.y = xy[2 * (size_t)i + 1];
            k.size = 7.f;
            k.angle = -1.f;
            k.response = resp[(size_t)i];
            k.octave = 0;
            k.class_id = -1;
        }
    }

  private:
    FastFeatureDetector(int t, bool n) : threshold_(t), nonmax_(n) {}
    int threshold_;
    bool nonmax_;
};

// cv::ORB as kitti_ba.cpp:128 builds it:  `Ptr<epivo::ORB> orb = epivo::ORB::create(10000, 1.2f, 8, 15, 0, 2, epivo::ORB::FAST_SCORE);`
// then `orb->detect(src, kp0, Mat());` and `orb->compute(src, kp0, desc0);` unchanged (kitti_ba.cpp:141,149).  detect() runs
// the whole detectAndCompute on the GPU and keeps the descriptors; compute() on the same image and the keypoints detect()
// returned hands them out (OpenCV's compute() on those keypoints produces exactly these rows).  Keypoints that did not come
// from this detector are refused: describing foreign keypoints is not part of the reference's path.
class ORB {
  public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };
    static std::shared_ptr<ORB> create(int nfeatures = 500, float scaleFactor = 1.2f, int nlevels = 8, int edgeThreshold = 31,
                                       int firstLevel = 0, int WTA_K = 2, int scoreType = HARRIS_SCORE, int patchSize = 31,
                                       int fastThreshold = 20) {
        if (firstLevel != 0 || WTA_K != 2 || scoreType != FAST_SCORE || patchSize != 31)
            throw std::invalid_argument("epivo::ORB: built for firstLevel 0, WTA_K 2, FAST_SCORE, patchSize 31 (kitti_ba.cpp:128)");
        return std::shared_ptr<ORB>(new ORB(nfeatures, scaleFactor, nlevels, edgeThreshold, fastThreshold));
    }
    template <typename MatT, typename KP>
    void detect(const MatT& image, std::vector<KP>& keypoints, const MatT& mask = MatT()) {
        typedef mat_traits<MatT> MT;
        if (!MT::empty(mask)) throw std::invalid_argument("ORB::detect: masks are not supported");
        run(image);
        keypoints.resize(kps_.size());
        for (size_t i = 0; i < kps_.size(); ++i) {
            KP& k = keypoints[i];
            k.pt.x = kps_[i].x;
            k.pt.y = kps_[i].y;
            k.size = kps_[i].size;
            k.angle = kps_[i].angle;
            k.response = kps_[i].response;
            k.octave = kps_[i].octave;
            k.class_id = kps_[i].class_id;
        }
    }
    template <typename MatT, typename KP>
    void compute(const MatT& image, std::vector<KP>& keypoints, MatT& descriptors) {
        typedef mat_traits<MatT> MT;
        if (!same_image(image)) run(image);
        bool same = keypoints.size() == kps_.size();
        for (size_t i = 0; same && i < kps_.size(); ++i)
            same = keypoints[i].pt.x == kps_[i].x && keypoints[i].pt.y == kps_[i].y && keypoints[i].octave == kps_[i].octave;
        if (!same) throw std::invalid_argument("ORB::compute: the keypoints are not the ones detect() returned for this image");
        descriptors = MT::create_u8((int)kps_.size(), 32);
        if (!kps_.empty()) memcpy(MT::bytes_mut(descriptors), desc_.data(), kps_.size() * 32);
    }
    template <typename MatT, typename KP>
    void detectAndCompute(const MatT& image, const MatT& mask, std::vector<KP>& keypoints, MatT& descriptors) {
        detect(image, keypoints, mask);
        compute(image, keypoints, descriptors);
    }

  private:
    ORB(int nf, float sf, int nl, int et, int ft) : nfeatures_(nf), scale_(sf), nlevels_(nl), edge_(et), fast_(ft), rows_(0), cols_(0), sum_(0) {}
    template <typename MatT>
    static uint64_t checksum(const MatT& image) {             // which image the cached result belongs to: a 64-bit
        typedef mat_traits<MatT> MT;                          // multiplicative hash over the pixels, 8 bytes per step
        const unsigned char* p = MT::bytes(image);
        const size_t n = (size_t)MT::rows(image) * MT::cols(image);
        uint64_t h = 1469598103934665603ull ^ n;
        size_t i = 0;
        for (; i + 8 <= n; i += 8) {
            uint64_t w;
            memcpy(&w, p + i, 8);
            h = (h ^ w) * 0x9E3779B97F4A7C15ull;
            h ^= h >> 29;
        }
        for (; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
        return h;
    }
    template <typename MatT>
    bool same_image(const MatT& image) const {
        typedef mat_traits<MatT> MT;
        return MT::rows(image) == rows_ && MT::cols(image) == cols_ && rows_ > 0 && checksum(image) == sum_;
    }
    template <typename MatT>
    void run(const MatT& image) {
        typedef mat_traits<MatT> MT;
        kps_.clear();
        desc_.clear();
        rows_ = MT::rows(image);
        cols_ = MT::cols(image);
        if (MT::empty(image)) return;
        sum_ = checksum(image);
        Context& ctx = default_context();
        int cap = nfeatures_ + nfeatures_ / 8 + 256;
        int32_t found = 0;
        for (;;) {                              // ties at a level's budget are all kept: a frame that overflows is run again
            kps_.assign((size_t)cap, epivo_keypoint());
            desc_.assign((size_t)cap * 32, 0);
            ctx.check(epivo_orb_detect_and_compute(ctx.get(), MT::bytes(image), 1, rows_, cols_, nfeatures_, scale_, nlevels_,
                                                   edge_, fast_, cap, kps_.data(), desc_.data(), &found));
            if (found <= cap) break;
            cap = found;
        }
        kps_.resize((size_t)found);
        desc_.resize((size_t)found * 32);
    }
    int nfeatures_;
    float scale_;
    int nlevels_, edge_, fast_;
    int rows_, cols_;
    uint64_t sum_;
    std::vector<epivo_keypoint> kps_;
    std::vector<unsigned char> desc_;
};

}  // namespace epivo

// void calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, status, err)                   kitti_E.cpp:79-84
// with OpenCV's defaults (21 x 21 window, 3 pyramid levels above the image, 30 iterations / 0.01).  status in {0,1}.
// err: OpenCV's window error (mean absolute difference at the final position), 0 for lost tracks.
template <typename MatT, typename Pt>
void calcOpticalFlowPyrLK(const MatT& prevImg, const MatT& nextImg, const std::vector<Pt>& prevPts, std::vector<Pt>& nextPts,
                          std::vector<unsigned char>& status, std::vector<float>& err) {
    typedef epivo::mat_traits<MatT> MT;
    if (MT::rows(prevImg) != MT::rows(nextImg) || MT::cols(prevImg) != MT::cols(nextImg))
        throw std::invalid_argument("calcOpticalFlowPyrLK: image sizes differ");
    const int n = (int)prevPts.size(), rows = MT::rows(prevImg), cols = MT::cols(prevImg);
    nextPts.assign((size_t)n, Pt());
    status.assign((size_t)n, 0);
    err.assign((size_t)n, 0.f);
    if (n == 0) return;
    epivo::Context& ctx = epivo::default_context();
    std::vector<unsigned char> frames(2 * (size_t)rows * cols);
    std::copy(MT::bytes(prevImg), MT::bytes(prevImg) + (size_t)rows * cols, frames.begin());
    std::copy(MT::bytes(nextImg), MT::bytes(nextImg) + (size_t)rows * cols, frames.begin() + (size_t)rows * cols);
    std::vector<float> p = epivo::detail::flatten(prevPts), q(2 * (size_t)n);
    const int32_t count = n;
    ctx.check(epivo_lk_track(ctx.get(), frames.data(), 2, rows, cols, p.data(), &count, n, 3, 30, 0.01, 1e-4, q.data(),
                             status.data(), err.data()));
    for (int i = 0; i < n; ++i) { nextPts[(size_t)i].x = q[2 * (size_t)i]; nextPts[(size_t)i].y = q[2 * (size_t)i + 1]; }
}

// void remap(src, dst, map1, map2, INTER_LINEAR)                                               euroc_E.cpp:170,174
// with the fixed-point maps initUndistortRectifyMap(..., map1.type() == 0 -> CV_16SC2, ...) produced at :105-113 (map1:
// rows x cols x 2 int16, map2: rows x cols uint16); BORDER_CONSTANT 0, as the call's defaults.  dst is created (8-bit,
// the size of the maps).  Only interpolation == INTER_LINEAR (1) is provided.
template <typename MatT>
void remap(const MatT& src, MatT& dst, const MatT& map1, const MatT& map2, int interpolation) {
    typedef epivo::mat_traits<MatT> MT;
    if (interpolation != 1) throw std::invalid_argument("remap: only INTER_LINEAR is provided");
    const int drows = MT::rows(map1), dcols = MT::cols(map1);
    if (MT::rows(map2) != drows || MT::cols(map2) != dcols) throw std::invalid_argument("remap: map sizes differ");
    dst = MT::create_u8(drows, dcols);
    if (drows == 0 || dcols == 0) return;
    epivo::Context& ctx = epivo::default_context();
    ctx.check(epivo_remap(ctx.get(), MT::bytes(src), 1, MT::rows(src), MT::cols(src),
                          reinterpret_cast<const int16_t*>(MT::bytes(map1)), reinterpret_cast<const uint16_t*>(MT::bytes(map2)),
                          drows, dcols, 0, MT::bytes_mut(dst)));
}
