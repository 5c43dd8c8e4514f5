// A miniature of test_jac_Rt_gen.cpp's synthetic LM demo and of the kitti_E.cpp:98-201 call
// sequence, written against include/epivo_shims.hpp with stand-in Matrix / Point2f types
// (Eigen and OpenCV headers are not in this image).  Built and run by tests/test_cpp_shims.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "epivo_shims.hpp"

struct Mat {                                   // the subset of Eigen::MatrixXd the shims touch
    int r = 0, c = 0;
    std::vector<double> v;
    Mat() {}
    Mat(int rows, int cols) : r(rows), c(cols), v((size_t)rows * cols, 0.0) {}
    int rows() const { return r; }
    int cols() const { return c; }
    double& operator()(int i, int j) { return v[(size_t)i * c + j]; }
    double operator()(int i, int j) const { return v[(size_t)i * c + j]; }
};
struct Point2f { float x, y; };

static double urand() { return (double)rand() / RAND_MAX; }

static Mat rot_xyz(double a, double b, double g) {
    Mat R(3, 3);
    double ca = cos(a), sa = sin(a), cb = cos(b), sb = sin(b), cg = cos(g), sg = sin(g);
    double Rx[9] = {1, 0, 0, 0, ca, -sa, 0, sa, ca}, Ry[9] = {cb, 0, sb, 0, 1, 0, -sb, 0, cb}, Rz[9] = {cg, -sg, 0, sg, cg, 0, 0, 0, 1};
    double t[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { t[i * 3 + j] = 0; for (int k = 0; k < 3; ++k) t[i * 3 + j] += Rx[i * 3 + k] * Ry[k * 3 + j]; }
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { double s = 0; for (int k = 0; k < 3; ++k) s += t[i * 3 + k] * Rz[k * 3 + j]; R(i, j) = s; }
    return R;
}

int main() {
    srand(7);
    epivo::Context ctx(0);
    // ---- Levenberg_Marquardt: one pair, noiseless correspondences, perturbed start (sequence.hpp:10-104)
    Mat R = rot_xyz(0.2, -0.1, 0.15);
    double t[3] = {0.4, -0.3, 1.5};
    const int N = 48;
    std::vector<Mat> T0s(1, Mat(4, 4)), pr(1, Mat(N, 3)), p_r(1, Mat(N, 3));
    for (int i = 0; i < N; ++i) {
        double X[3] = {20 * (urand() - 0.5), 20 * (urand() - 0.5), 15 + 20 * urand()}, Y[3];
        for (int a = 0; a < 3; ++a) Y[a] = R(a, 0) * X[0] + R(a, 1) * X[1] + R(a, 2) * X[2] + t[a];
        for (int a = 0; a < 3; ++a) { pr[0](i, a) = X[a] / X[2]; p_r[0](i, a) = Y[a] / Y[2]; }
    }
    Mat Rn = rot_xyz(0.2 + 0.03, -0.1 - 0.02, 0.15 + 0.02);
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T0s[0](i, j) = Rn(i, j); T0s[0](i, 3) = t[i] + 0.05 * (urand() - 0.5); }
    T0s[0](3, 3) = 1.0;
    std::vector<std::pair<int, int> > reps(1, std::make_pair(0, 0));
    std::vector<double> wreps(1, 1.0);
    LM_res res;
    epivo::Levenberg_Marquardt(ctx, 1, 1e-8, reps, wreps, 1e-2, T0s, pr, p_r, res, /*huber_delta=*/1.0, /*max_iters=*/60);
    double dR = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dR += (T0s[0](i, j) - R(i, j)) * (T0s[0](i, j) - R(i, j));
    double ratio[3] = {t[0] / T0s[0](0, 3), t[1] / T0s[0](1, 3), t[2] / T0s[0](2, 3)};
    printf("LM |R-R0| %.3e r_norm %.3e lambda %.3e t ratios %.6f %.6f %.6f\n", sqrt(dR), res.r_norm, res.lambda, ratio[0], ratio[1], ratio[2]);
    if (sqrt(dR) > 1e-6 || fabs(ratio[0] - ratio[2]) > 1e-5 || fabs(ratio[1] - ratio[2]) > 1e-5) return 1;

    // ---- findEssentialMat + recoverPose on pixels of the same scene (kitti_E.cpp:98-120)
    const double cam[9] = {718.856, 0, 607.1928, 0, 718.856, 185.2157, 0, 0, 1};
    std::vector<Point2f> p0(N), p1(N);
    for (int i = 0; i < N; ++i) {
        p0[i].x = (float)(cam[0] * pr[0](i, 0) + cam[2]); p0[i].y = (float)(cam[4] * pr[0](i, 1) + cam[5]);
        p1[i].x = (float)(cam[0] * p_r[0](i, 0) + cam[2]); p1[i].y = (float)(cam[4] * p_r[0](i, 1) + cam[5]);
    }
    double E[9], Rr[9], tr[3];
    std::vector<unsigned char> mask_ess, rec_mask;
    if (!epivo::findEssentialMat(ctx, p0, p1, cam, epivo::RANSAC, 0.99, 1.0, E, mask_ess)) return 2;
    int ninl = 0;
    for (size_t i = 0; i < mask_ess.size(); ++i) ninl += (int)mask_ess[i] == 1;
    int good = epivo::recoverPose(ctx, E, p0, p1, cam, Rr, tr, rec_mask);
    double dRr = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) dRr += (Rr[i * 3 + j] - R(i, j)) * (Rr[i * 3 + j] - R(i, j));
    printf("E inliers %d/%d  recoverPose good %d  |R-Rgt| %.3e\n", ninl, N, good, sqrt(dRr));
    if (ninl < N - 2 || good < N / 2 || sqrt(dRr) > 1e-3) return 3;

    // ---- BFMatcher(NORM_HAMMING2, true) on identical descriptor sets: the identity matching
    std::vector<uint8_t> d(64 * 32);
    for (size_t i = 0; i < d.size(); ++i) d[i] = (uint8_t)(rand() & 255);
    std::vector<epivo::DMatch> matches;
    epivo::BFMatcher(ctx, epivo::NORM_HAMMING2, true).match(d.data(), 64, d.data(), 64, 32, matches);
    if (matches.size() != 64) return 4;
    for (int i = 0; i < 64; ++i) if (matches[i].queryIdx != i || matches[i].trainIdx != i || matches[i].distance != 0.f) return 5;
    printf("shims ok\n");
    return 0;
}
