// The reference's hot-path call lines, LITERALLY, compiled against include/epivo_dropin.hpp.
//
// OpenCV and Eigen are not in this image, so this file first defines stand-ins with the same member API the
// drop-in touches -- a `cv` namespace (Mat with rows / cols / data / at<T>() / depth() / empty(), Mat_<T> with the
// comma initialiser of kitti_E.cpp:38, Point2f, DMatch, the LMEDS / RANSAC / NORM_HAMMING2 / CV_32F / CV_64F
// constants, cv2eigen) and the Eigen stand-in of oracle/ref_shim -- and then, like the drivers, says
// `using namespace cv; using namespace std; using namespace Eigen;` and includes the drop-in header.  The lines
// marked [verbatim] are the reference's call lines, character for character:
//   kitti_E.cpp:38-40 (cam), :98-104 (findEssentialMat), :120 (recoverPose), :196 (7-argument LM),
//   kitti_ba.cpp:602,641 (BFMatcher: class name qualified with epivo::, the one edit), :702, :715, :881 (9-argument LM).
// Around them the loop body of kitti_E.cpp:96-201 is replayed on a synthetic pair.  Built by
// tests/test_cpp_shims.py (compiles everywhere; runs on a GPU box).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <utility>
#include <vector>

#include <Eigen/Dense>          // oracle/ref_shim stand-in

#define OPENCV_CORE_MAT_HPP     // what <opencv2/core/mat.hpp> defines: enables the cv::Mat traits of the drop-in
#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_16SC2 11       /* CV_MAKETYPE(CV_16S, 2) */
#define CV_16UC1 2
#define INTER_LINEAR 1
typedef unsigned char uchar;
namespace cv {
enum { LMEDS = 4, RANSAC = 8, NORM_HAMMING = 6, NORM_HAMMING2 = 7 };
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct DMatch { int queryIdx, trainIdx, imgIdx; float distance; };
struct KeyPoint {
    Point2f pt; float size, angle, response; int octave, class_id;
    static void convert(const std::vector<KeyPoint>& k, std::vector<Point2f>& p) { p.clear(); for (size_t i = 0; i < k.size(); ++i) p.push_back(k[i].pt); }
};
template <typename T> struct Ptr : std::shared_ptr<T> { Ptr() {} Ptr(const std::shared_ptr<T>& o) : std::shared_ptr<T>(o) {} };
class Mat {
  public:
    int rows, cols;
    uchar* data;
    Mat() : rows(0), cols(0), data(0), type_(CV_8U) {}
    Mat(int r, int c, int type) : rows(r), cols(c), type_(type), buf_((size_t)r * c * esz(type), 0) { data = buf_.data(); }
    Mat(const Mat& o) : rows(o.rows), cols(o.cols), type_(o.type_), buf_(o.buf_) { data = buf_.empty() ? 0 : buf_.data(); }
    Mat& operator=(const Mat& o) { rows = o.rows; cols = o.cols; type_ = o.type_; buf_ = o.buf_; data = buf_.empty() ? 0 : buf_.data(); return *this; }
    int depth() const { return type_; }
    bool empty() const { return rows == 0 || cols == 0; }
    template <typename T> T& at(int i, int j) { return reinterpret_cast<T*>(data)[(size_t)i * cols + j]; }
    template <typename T> const T& at(int i, int j) const { return reinterpret_cast<const T*>(data)[(size_t)i * cols + j]; }
  private:
    static size_t esz(int t) { return t == CV_64F ? 8 : t == CV_32F || t == CV_16SC2 ? 4 : t == CV_16UC1 ? 2 : 1; }
    int type_;
    std::vector<uchar> buf_;
};
template <typename T> struct MatCommaInit {
    Mat m; int k;
    MatCommaInit& operator,(double v) { m.at<T>(k / m.cols, k % m.cols) = (T)v; ++k; return *this; }
    operator Mat() const { return m; }
};
template <typename T> class Mat_ {
  public:
    Mat_(int r, int c) : r_(r), c_(c) {}
    MatCommaInit<T> operator<<(double v) const {
        MatCommaInit<T> ci = {Mat(r_, c_, sizeof(T) == 4 ? CV_32F : CV_64F), 0};
        ci, v;
        return ci;
    }
  private:
    int r_, c_;
};
inline void cv2eigen(const Mat& src, Eigen::MatrixXd& dst) {
    dst = Eigen::MatrixXd(src.rows, src.cols);
    for (int i = 0; i < src.rows; ++i) for (int j = 0; j < src.cols; ++j) dst(i, j) = src.at<double>(i, j);
}
}  // namespace cv

#include "epivo_dropin.hpp"     // the ONE added include

using namespace cv;
using namespace std;
using namespace Eigen;

static double urand() { return (double)rand() / RAND_MAX; }

int main() {
    srand(11);
    // [verbatim] kitti_E.cpp:38-40
    Mat cam = (Mat_<float>(3,3) << 718.8560, 0.0, 607.1928,
                                    0.0, 718.8560, 185.2157,
                                    0.0, 0.0,      1.0);
    MatrixXd cam_(3, 3);
    cam_ << 718.8560, 0.0,      607.1928,
            0.0,      718.8560, 185.2157,
            0.0,      0.0,      1.0;
    cam_ = cam_.inverse();

    // a synthetic pair: 300 landmarks seen from two poses, 0.3 px noise, 20 % gross outliers
    const double ang[3] = {0.012, -0.02, 0.008}, tgt[3] = {0.03, -0.02, -1.0};
    MatrixXd Rx(3, 3), Ry(3, 3), Rz(3, 3);
    Rx << 1, 0, 0, 0, cos(ang[0]), -sin(ang[0]), 0, sin(ang[0]), cos(ang[0]);
    Ry << cos(ang[1]), 0, sin(ang[1]), 0, 1, 0, -sin(ang[1]), 0, cos(ang[1]);
    Rz << cos(ang[2]), -sin(ang[2]), 0, sin(ang[2]), cos(ang[2]), 0, 0, 0, 1;
    const MatrixXd Rgt = Rx * Ry * Rz;
    vector<Point2f> _cpt0, _cpt1_;
    for (int i = 0; i < 300; ++i) {
        const double X[3] = {30 * (urand() - 0.5), 8 * (urand() - 0.5), 10 + 40 * urand()};
        double Y[3];
        for (int a = 0; a < 3; ++a) Y[a] = Rgt(a, 0) * X[0] + Rgt(a, 1) * X[1] + Rgt(a, 2) * X[2] + tgt[a];
        Point2f a((float)(718.856 * X[0] / X[2] + 607.1928), (float)(718.856 * X[1] / X[2] + 185.2157));
        Point2f b((float)(718.856 * Y[0] / Y[2] + 607.1928 + 0.6 * (urand() - 0.5)),
                  (float)(718.856 * Y[1] / Y[2] + 185.2157 + 0.6 * (urand() - 0.5)));
        if (i % 5 == 4) b = Point2f((float)(1241 * urand()), (float)(376 * urand()));
        _cpt0.push_back(a);
        _cpt1_.push_back(b);
    }

    vector<uchar> mask_ess;
    // [verbatim] kitti_E.cpp:98-104
    Mat ess = findEssentialMat(_cpt0,
                               _cpt1_,
                               cam,
                               LMEDS,
                               0.99,
                               0.01,
                               mask_ess);
    if (ess.rows != 3 || ess.cols != 3 || ess.depth() != CV_64F || mask_ess.size() != _cpt0.size()) return 1;
    vector<Point2f> cpt0, cpt1_;
    for (int j = 0; j < (int)mask_ess.size(); j++) {
        if ((int)mask_ess[j] != 0 && (int)mask_ess[j] != 1) return 2;          // {0,1}, as cv2 (:108 tests == 1)
        if ((int)mask_ess[j] == 1) { cpt0.push_back(_cpt0[j]); cpt1_.push_back(_cpt1_[j]); }
    }
    Mat rot, tr;
    vector<uchar> rec_mask;
    // [verbatim] kitti_E.cpp:120
    recoverPose(ess, cpt0, cpt1_, cam, rot, tr, rec_mask); // mask_ess
    MatrixXd erot(3, 3), etr(3, 1);
    cv2eigen(rot, erot);
    cv2eigen(tr, etr);
    int n255 = 0;
    for (size_t j = 0; j < rec_mask.size(); ++j) {
        if (rec_mask[j] != 0 && rec_mask[j] != 255) return 3;                  // {0,255} (:177 tests == 255)
        n255 += rec_mask[j] == 255;
    }
    printf("findEssentialMat inliers %zu / %zu, recoverPose good %d, |R - Rgt| %.3e, trace %.6f\n", cpt0.size(),
           _cpt0.size(), n255, (erot - Rgt).norm(), erot.trace());
    if (cpt0.size() < 200 || n255 < 150 || (erot - Rgt).norm() > 5e-3 || rec_mask.size() != cpt0.size()) return 4;

    vector<pair<int, int> > reps;
    reps.push_back(make_pair(0, 0));
    vector<MatrixXd> T0s;
    MatrixXd T0_0 = MatrixXd::Identity(4, 4);
    T0_0.block<3, 3>(0, 0) = erot;
    T0_0.block<3, 1>(0, 3) = etr;
    T0s.push_back(T0_0);
    vector<MatrixXd> bT0s(T0s);
    vector<MatrixXd> pr, p_r;
    int N = 48;
    MatrixXd pr_(N, 3), p_r_(N, 3);
    for (int j = 0; j < N; j++) {                                              // kitti_E.cpp:173-186 (first N inliers)
        MatrixXd a(3, 1), b(3, 1);
        a << cpt0[j].x, cpt0[j].y, 1.0;
        b << cpt1_[j].x, cpt1_[j].y, 1.0;
        pr_.row(j) = (cam_ * a).transpose();
        p_r_.row(j) = (cam_ * b).transpose();
    }
    pr.push_back(pr_);
    p_r.push_back(p_r_);
    double uncert;
    // [verbatim] kitti_E.cpp:196
    uncert = Levenberg_Marquardt(1, 1e-8, reps, 1e-2, T0s, pr, p_r);
    printf("7-argument LM: uncert (r_norm) %.6e, |T - T0| %.3e\n", uncert, (T0s[0] - bT0s[0]).norm());
    if (!(uncert == uncert) || !(uncert < 1e-3) || (T0s[0] - bT0s[0]).norm() == 0.0) return 5;

    // kitti_ba.cpp:876-882: the 9-argument form on a two-zeta chain (frames 0 -> 1 -> 2, the second step = the first)
    {
        vector<pair<int, int> > reps;
        vector<double> wreps;
        reps.push_back(make_pair(0, 0)); wreps.push_back(1.0);
        reps.push_back(make_pair(1, 1)); wreps.push_back(1.0);
        vector<MatrixXd> T0s, pr, p_r;
        T0s.push_back(bT0s[0]); T0s.push_back(bT0s[0]);
        pr.push_back(pr_); pr.push_back(pr_);
        p_r.push_back(p_r_); p_r.push_back(p_r_);
        int nzeta = 2;
        LM_res lm_res;
        // [verbatim] kitti_ba.cpp:881
            Levenberg_Marquardt(nzeta, 1e-8, reps, wreps, 1e-2, T0s, pr, p_r, lm_res);
        printf("9-argument LM: H_norm %.3e r_norm %.6e lambda %.3e\n", lm_res.H_norm, lm_res.r_norm, lm_res.lambda);
        if (!(lm_res.r_norm < 1e-3) || !(lm_res.H_norm > 0) || !(lm_res.lambda > 0)) return 6;
        if (fabs(lm_res.r_norm - sqrt(2.0) * uncert) > 1e-6 * uncert + 1e-12) return 7;   // two copies of the same problem
    }

    // kitti_ba.cpp:602,641,702,715: descriptors -> matches -> E (LMEDS .99 .1) -> pose
    {
        Mat desc0(200, 32, CV_8U), desc1(200, 32, CV_8U);
        for (int i = 0; i < 200 * 32; ++i) desc0.data[i] = (uchar)(rand() & 255);
        for (int i = 0; i < 200; ++i) {                                       // frame 1 = frame 0 reversed, 8 % bits flipped
            memcpy(desc1.data + (size_t)(199 - i) * 32, desc0.data + (size_t)i * 32, 32);
            for (int b = 0; b < 20; ++b) desc1.data[(size_t)(199 - i) * 32 + (rand() % 32)] ^= (uchar)(1 << (rand() & 7));
        }
        // [verbatim, class name qualified] kitti_ba.cpp:602
        epivo::BFMatcher matcher(NORM_HAMMING2, true); // , true
        vector<DMatch> matches;
        // [verbatim] kitti_ba.cpp:641
            matcher.match(desc0, desc1, matches);
        if (matches.size() != 200) return 8;
        for (int i = 0; i < 200; ++i)
            if (matches[i].queryIdx != i || matches[i].trainIdx != 199 - i || matches[i].imgIdx != 0 || matches[i].distance > 40.f) return 9;
        vector<Point2f>& _cpt1 = _cpt1_;
        vector<uchar> mask_ess;
        vector<Point2f> cpt0, cpt1;
        Mat rot, tr;
        vector<uchar> rec_mask;
            if(_cpt0.size() >= 8){
                // [verbatim] kitti_ba.cpp:702
                Mat ess = findEssentialMat(_cpt0, _cpt1, cam, LMEDS, 0.99, 0.1, mask_ess);
                for(int k = 0; k < (int)mask_ess.size(); k++){
                    if((int)mask_ess[k] == 1){
                        cpt0.push_back(_cpt0[k]);
                        cpt1.push_back(_cpt1[k]);
                    }
                }
                // [verbatim] kitti_ba.cpp:715
                recoverPose(ess, cpt0, cpt1, cam, rot, tr, rec_mask);
            }
        if (rot.rows != 3 || tr.rows != 3 || tr.cols != 1 || cpt0.size() < 200) return 10;
    }
    // kitti_E.cpp:66-95: detector and tracker on a synthetic image pair (smooth texture shifted by (2, -1) pixels)
    {
        const int rows = 120, cols = 200;
        Mat big(rows + 8, cols + 8, CV_8U);
        for (int y = 0; y < big.rows; ++y)
            for (int x = 0; x < big.cols; ++x)
                big.at<uchar>(y, x) = (uchar)(127.5 + 60 * sin(0.31 * x) * cos(0.27 * y) + 50 * sin(0.11 * x * y * 0.05) + 10 * ((x * 7 + y * 13) % 5));
        Mat src(rows, cols, CV_8U), tgt(rows, cols, CV_8U);
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) { src.at<uchar>(y, x) = big.at<uchar>(y + 4, x + 4); tgt.at<uchar>(y, x) = big.at<uchar>(y + 5, x + 2); }
        vector<KeyPoint> kp0, kp_; // kp1,
        // [verbatim, class name qualified] kitti_E.cpp:70
        Ptr<epivo::FastFeatureDetector> detector = epivo::FastFeatureDetector::create(40);
        // [verbatim] kitti_E.cpp:73
        detector->detect(src, kp0, Mat());
        vector<Point2f> pt0, pt1_;
        // [verbatim] kitti_E.cpp:77
        cv::KeyPoint::convert(kp0, pt0);
        vector<uchar> status;
        vector<float> err;
        // [verbatim] kitti_E.cpp:79-84
        calcOpticalFlowPyrLK(src,
                             tgt,
                             pt0,
                             pt1_,
                             status,
                             err);
        int tracked = 0, close = 0;
        for (size_t j = 0; j < status.size(); j++) {
            if ((int)status[j] == 1) {                                         // kitti_E.cpp:88
                ++tracked;
                close += fabs(pt1_[j].x - pt0[j].x - 2.0) < 0.3 && fabs(pt1_[j].y - pt0[j].y + 1.0) < 0.3;
            }
        }
        // euroc_E.cpp:169-174: undistortion remap with fixed-point maps (here: a half-pixel shift to the right and down)
        Mat map1(rows, cols, CV_16SC2), map2(rows, cols, CV_16UC1), src_;
        for (int y = 0; y < rows; ++y)
            for (int x = 0; x < cols; ++x) {
                reinterpret_cast<short*>(map1.data)[2 * ((size_t)y * cols + x)] = (short)x;
                reinterpret_cast<short*>(map1.data)[2 * ((size_t)y * cols + x) + 1] = (short)y;
                reinterpret_cast<unsigned short*>(map2.data)[(size_t)y * cols + x] = (unsigned short)(16 * 32 + 16);
            }
        // [verbatim] euroc_E.cpp:170
        remap(src, src_, map1, map2, INTER_LINEAR);
        if (src_.rows != rows || src_.cols != cols) return 13;
        {
            const int y = 50, x = 70;            // mean of the 2 x 2 neighbourhood, rounded half up
            const int want = (src.at<uchar>(y, x) + src.at<uchar>(y, x + 1) + src.at<uchar>(y + 1, x) + src.at<uchar>(y + 1, x + 1) + 2) >> 2;
            if (src_.at<uchar>(y, x) != want || src_.at<uchar>(rows - 1, cols - 1) != ((src.at<uchar>(rows - 1, cols - 1) + 2) >> 2)) return 14;
        }
        printf("FAST(40): %zu corners, LK tracked %d, %d within 0.3 px of the true shift\n", kp0.size(), tracked, close);
        if (kp0.size() < 20 || pt1_.size() != pt0.size() || tracked < (int)kp0.size() * 8 / 10 || close < tracked * 8 / 10) return 11;
        if (kp0[0].size != 7.f || kp0[0].angle != -1.f || kp0[0].response < 40.f) return 12;
    }
    // kitti_ba.cpp:114-156 (extract_good_kp): ORB keypoints and descriptors of two frames, matched as really_robust_ass does
    {
        const int rows = 160, cols = 240;
        Mat frames[2] = {Mat(rows, cols, CV_8U), Mat(rows, cols, CV_8U)};
        unsigned lcg = 12345u;
        std::vector<int> blocks((rows / 8 + 2) * (cols / 8 + 2));
        for (size_t i = 0; i < blocks.size(); ++i) { lcg = lcg * 1664525u + 1013904223u; blocks[i] = 40 + (lcg >> 24) % 180; }
        for (int f = 0; f < 2; ++f)
            for (int y = 0; y < rows; ++y)
                for (int x = 0; x < cols; ++x)             // a blocky texture (corners everywhere), frame 1 shifted by (3, 1)
                    frames[f].at<uchar>(y, x) = (uchar)(blocks[((y + f) / 8) * (cols / 8 + 2) + (x + 3 * f) / 8] + ((x * 5 + y * 3) % 7));
        vector<vector<Point2f> > key_points;
        vector<Mat> descs;
        // [verbatim, class name qualified] kitti_ba.cpp:128
        Ptr<epivo::ORB> orb = epivo::ORB::create(10000, 1.2f, 8, 15, 0, 2, epivo::ORB::FAST_SCORE);
        for (int i = 0; i < 2; i++) {
            Mat src = frames[i];
            // [verbatim] kitti_ba.cpp:140-153
            vector<KeyPoint> kp0;
            orb->detect(src, kp0, Mat());
            //detector.detect(src, kp0);

            vector<Point2f> pt0;
            cv::KeyPoint::convert(kp0, pt0);

            Mat desc0;
            //extractor.compute(src, kp0, desc0);
            orb->compute(src, kp0, desc0);

            key_points.push_back(pt0);
            descs.push_back(desc0);
            if (kp0.size() < 50 || desc0.rows != (int)kp0.size() || desc0.cols != 32) return 15;
            if (kp0[0].octave != 0 || kp0[0].size != 31.f || kp0[0].angle < 0.f || kp0[0].angle >= 360.f || kp0[0].response < 20.f) return 16;
        }
        // [verbatim] kitti_ba.cpp:602,641
        epivo::BFMatcher matcher(NORM_HAMMING2, true);
        vector<DMatch> matches;
        matcher.match(descs[0], descs[1], matches);
        int good = 0;
        for (size_t m = 0; m < matches.size(); ++m) {
            const Point2f a = key_points[0][matches[m].queryIdx], b = key_points[1][matches[m].trainIdx];
            good += fabs(a.x - b.x - 3.0) < 2.5 && fabs(a.y - b.y - 1.0) < 2.5;
        }
        printf("ORB: %zu / %zu keypoints, %zu mutual matches, %d on the true shift\n", key_points[0].size(), key_points[1].size(),
               matches.size(), good);
        if (matches.size() < 30 || good < (int)matches.size() / 2) return 17;
    }
    printf("dropin ok\n");
    return 0;
}
