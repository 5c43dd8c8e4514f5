// ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// Compiles the reference's own LM translation unit, UNMODIFIED, where it lies
// (/root/reference/jac_Rt_gen_.cpp, which in turn includes test_jac_Rt_gen.hpp and sequence.hpp),
// against the Eigen / Sophus stand-ins in oracle/ref_shim/, and exports plain-C entry points so
// that the numpy restatement (oracle/oracle.py), the C restatement (oracle/lm_c.c) and the CUDA
// kernels can be compared with the reference's code itself.  Outputs go to oracle/_ref/ only
// (git-ignored); no reference source is copied into this repository.
//
// What this file adds around the #include:
//   * `struct LM_res` -- used at jac_Rt_gen_.cpp:295,473-475 and kitti_ba.cpp:876 but defined
//     nowhere in the reference; fields taken from the three assignments at :473-475.
//   * REF_TU selects the translation unit: jac_Rt_gen_.cpp (default; huber_delta = 1e-5, LM as a
//     function) or test_jac_Rt_gen.cpp (REF_TU_DEMO; huber_delta = 1.0, forward-only
//     RepJacobian, LM inlined in main(), which is renamed and seeded instead of time(0)).
//   * REF_PREFIX prefixes the exported names so several builds can be loaded side by side.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>
#include <utility>
#include <vector>
#include <Eigen/Dense>

struct LM_res { double H_norm, r_norm, lambda; };

#ifdef REF_TU_DEMO
static unsigned ref_demo_seed = 0;
#define time(x) ((time_t)ref_demo_seed)
#define main ref_demo_main
#include "test_jac_Rt_gen.cpp"
#undef main
#undef time
#else
#include REF_TU_FILE
#endif

#define REF_CAT2(a, b) a##b
#define REF_CAT(a, b) REF_CAT2(a, b)
#define REF_NAME(n) REF_CAT(REF_PREFIX, n)

namespace {
MatrixXd from_rows(const double* a, int r, int c) {          // row-major C array -> MatrixXd
    MatrixXd m(r, c);
    for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m(i, j) = a[i * c + j];
    return m;
}
void to_rows(const MatrixXd& m, double* a) {
    for (int i = 0; i < m.rows(); ++i) for (int j = 0; j < m.cols(); ++j) a[i * m.cols() + j] = m(i, j);
}
void fill_memo(const std::vector<MatrixXd>& T0s) {            // as jac_Rt_gen_.cpp:328-335
    const int n_zeta = (int)T0s.size();
    for (int j = 0; j < n_zeta; j++) {
        MatrixXd sT = T0s[j];
        T0_mem[j][j] = T0s[j];
        for (int k = j + 1; k < n_zeta; k++) { sT = T0s[k] * sT; T0_mem[j][k] = sT; }
    }
}
struct MuteCout {                                             // the reference prints diagnostics to cout
    std::streambuf* old; std::ostringstream sink;
    MuteCout() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~MuteCout() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

double REF_NAME(huber_delta)(void) { return huber_delta; }

// res(): jac_Rt_gen_.cpp:212-259.  R0 3x3 row-major, p / p_ N x 3 row-major, r[N] out.
int REF_NAME(res)(const double* R0, const double* t0, const double* p, const double* p_, int N, double* r) {
    MuteCout mute;
    MatrixXd R = from_rows(R0, 3, 3), P = from_rows(p, N, 3), P_ = from_rows(p_, N, 3);
    VectorXd t = from_rows(t0, 3, 1);
    MatrixXd out = MatrixXd::Zero(N, 1);
    res(R, t, P, P_, out);
    to_rows(out, r);
    return 0;
}

// Dr_Deps(): jac_Rt_gen_.cpp:23-209.  Tl0 / Tr0 4x4 row-major, J[N*6] row-major out.
int REF_NAME(dr_deps)(const double* Tl0, const double* Tr0, const double* p, const double* p_, int N,
                      int reverse, double* J) {
    MuteCout mute;
    MatrixXd Tl = from_rows(Tl0, 4, 4), Tr = from_rows(Tr0, 4, 4), P = from_rows(p, N, 3), P_ = from_rows(p_, N, 3);
    MatrixXd out = MatrixXd::Zero(N, 6);
#ifdef REF_TU_DEMO
    if (reverse) return -1;                                   // the demo's Dr_Deps has no `reverse`
    Dr_Deps(Tl, Tr, P, P_, out);
#else
    Dr_Deps(Tl, Tr, P, P_, reverse != 0, out);
#endif
    to_rows(out, J);
    return 0;
}

// RepJacobian(z, s, t).compute(): jac_Rt_gen_.cpp:262-284, with the memo built as at :328-335.
int REF_NAME(rep_jacobian)(int n_zeta, const double* T0s, int z, int s, int t, const double* p, const double* p_,
                           int N, double* J) {
    MuteCout mute;
    std::vector<MatrixXd> T;
    for (int k = 0; k < n_zeta; ++k) T.push_back(from_rows(T0s + 16 * k, 4, 4));
    fill_memo(T);
    MatrixXd P = from_rows(p, N, 3), P_ = from_rows(p_, N, 3);
    MatrixXd out = MatrixXd::Zero(N, 6);
    RepJacobian Jr(z, s, t);
    Jr.compute(P, P_, T, out);
    to_rows(out, J);
    return 0;
}

// SE3::exp (jac_Rt_gen_.cpp:419) through the Sophus stand-in: a = (upsilon, omega) -> 4x4 row-major.
void REF_NAME(se3_exp)(const double* a, double* T) {
    MatrixXd v = from_rows(a, 6, 1);
    to_rows(Sophus::SE3<double>::exp(v).matrix(), T);
}

// gen_scene_sequence(): sequence.hpp:106-159, seeded with srand(seed) instead of time(0).
// Ts / T0s: n_zeta x 16; Xr / pr / p_r: n_rep x N x 3, all row-major.
int REF_NAME(gen_scene_sequence)(unsigned seed, int N, int n_zeta, const int* reps, int n_rep, double* Ts,
                                 double* T0s, double* Xr, double* pr, double* p_r) {
    MuteCout mute;
    std::srand(seed);
    std::vector<std::pair<int, int> > rp;
    for (int j = 0; j < n_rep; ++j) rp.push_back(std::make_pair(reps[2 * j], reps[2 * j + 1]));
    std::vector<MatrixXd> vT, vT0, vX, vp, vp_;
    gen_scene_sequence(N, n_zeta, rp, vT, vT0, vX, vp, vp_);
    for (int k = 0; k < n_zeta; ++k) { to_rows(vT[k], Ts + 16 * k); to_rows(vT0[k], T0s + 16 * k); }
    for (int j = 0; j < n_rep; ++j) {
        to_rows(vX[j], Xr + (size_t)j * N * 3);
        to_rows(vp[j], pr + (size_t)j * N * 3);
        to_rows(vp_[j], p_r + (size_t)j * N * 3);
    }
    return 0;
}

#ifndef REF_TU_DEMO
// Levenberg_Marquardt(): jac_Rt_gen_.cpp:287-478, the 9-argument form.  T0s updated in place;
// out = {H_norm, r_norm, lambda}; nan_break reports whether "delta has Nan" was printed (:407-410);
// trace (optional, 2 x trace_cap doubles) receives (size, value) of the LM's norm() calls, see below.
int REF_NAME(lm)(int n_zeta, double epsilon, const int* reps, const double* wreps, int n_rep, double lambda0,
                 double* T0s, const double* pr, const double* p_r, int N, double* out, int* nan_break,
                 double* trace, int trace_cap, int* n_trace) {
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    std::vector<std::pair<int, int> > rp;
    std::vector<double> w(wreps, wreps + n_rep);
    for (int j = 0; j < n_rep; ++j) rp.push_back(std::make_pair(reps[2 * j], reps[2 * j + 1]));
    std::vector<MatrixXd> T, vp, vp_;
    for (int k = 0; k < n_zeta; ++k) T.push_back(from_rows(T0s + 16 * k, 4, 4));
    for (int j = 0; j < n_rep; ++j) {
        vp.push_back(from_rows(pr + (size_t)j * N * 3, N, 3));
        vp_.push_back(from_rows(p_r + (size_t)j * N * 3, N, 3));
    }
    LM_res lr = {0, 0, 0};
    std::vector<double> tr;
    if (trace) Eigen::epivo_norm_trace = &tr;
    int rc = Levenberg_Marquardt(n_zeta, epsilon, rp, w, lambda0, T, vp, vp_, lr);
    Eigen::epivo_norm_trace = nullptr;
    std::cout.rdbuf(old);
    if (trace) {
        // norm() calls of size > 3 inside the LM, in order: per iteration |delta| (size 6 n_zeta) and, unless
        // the loop broke on it, the candidate |r0| (size n_rep N); after the loop |H| and |r0| (:473-474).
        int n = 0;
        for (size_t k = 0; k + 1 < tr.size() && n < trace_cap; k += 2) { trace[2 * n] = tr[k]; trace[2 * n + 1] = tr[k + 1]; ++n; }
        if (n_trace) *n_trace = n;
    }
    for (int k = 0; k < n_zeta; ++k) to_rows(T[k], T0s + 16 * k);
    out[0] = lr.H_norm; out[1] = lr.r_norm; out[2] = lr.lambda;
    if (nan_break) *nan_break = sink.str().find("delta has Nan") != std::string::npos;
    return rc;
}
#else
// The convergence demo test_jac_Rt_gen.cpp:279-513 (n_zeta = 10, N = 15, reps (i,i),(0,i), 60 LM
// iterations at huber_delta = 1.0), seeded.  It prints ||R - R0|| and the t / t0 ratios per zeta and
// writes est.pose / gt.pose into the current directory; the text it prints is returned in `log`.
int REF_NAME(demo)(unsigned seed, char* log, int log_cap) {
    ref_demo_seed = seed;
    std::streambuf* old = std::cout.rdbuf();
    std::ostringstream sink;
    std::cout.rdbuf(sink.rdbuf());
    ref_demo_main();
    std::cout.rdbuf(old);
    std::string s = sink.str();
    if (log && log_cap > 0) { std::strncpy(log, s.c_str(), log_cap - 1); log[log_cap - 1] = 0; }
    return (int)s.size();
}
#endif

}  // extern "C"
