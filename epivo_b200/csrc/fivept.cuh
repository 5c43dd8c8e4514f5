// K2: Nister 5-point minimal solver, one hypothesis per thread, FP64.
//
// Restates the algorithm of cv::findEssentialMat's EMEstimatorCallback::runKernel
// (OpenCV modules/calib3d/src/five-point.cpp; called from kitti.cpp:98, kitti_E.cpp:98,
// euroc_E.cpp:202, kitti_ba.cpp:232,308,702):
//   1. 5x9 epipolar system (row-major E, x2' E x1 = 0) -> 4-D null space
//      (OpenCV: full SVD; here: Householder QR of the transpose -- any orthonormal basis of
//      the same null space yields the same set of essential matrices),
//   2. the ten cubic constraints -> Nister's 10x20 matrix (fivept_gen.cuh),
//   3. Gauss-Jordan with partial pivoting on the left 10x10 block (OpenCV: inv() * right),
//   4. B(z) = {rows 4,6,8} - z {rows 5,7,9}, det B(z) = degree-10 polynomial,
//   5. all complex roots by the Durand-Kerner iteration of cv::solvePoly (same start
//      values (1+i)^k, same Gauss-Seidel sweep, 300 sweeps at most; stops early once every
//      update of a sweep is below 1e-13 of its root -- OpenCV keeps sweeping at round-off level),
//   6. per root with |imag| <= 1e-10: (x, y) from the null vector of B(z) (skipped if its
//      third component is < 1e-10 in magnitude), E = x E0 + y E1 + z E2 + E3, normalised,
//   7. (not in OpenCV) each solution is refined by Gauss-Newton on the ten constraints inside the
//      null space (refine_essential below), which removes the basis-dependent loss of accuracy.
#pragma once
#include "fivept_gen.cuh"

namespace fivept {

// Tuning builds only (-DEPV_PROFILE_SOLVE): per-phase clock totals of solve(), read back by
// epivo_debug_solve_profile().  [0] null space + constraints [1] Gauss-Jordan [2] polynomial
// [3] Durand-Kerner [4] roots -> E + refinement [5] DK sweeps [6] solves [7] models
#ifdef EPV_PROFILE_SOLVE
__device__ unsigned long long g_solve_prof[8];
#define EPV_PROF_BEGIN long long _pt = clock64()
#define EPV_PROF(i)                                                       \
    do {                                                                  \
        const long long _n = clock64();                                   \
        atomicAdd(&g_solve_prof[i], (unsigned long long)(_n - _pt));      \
        _pt = _n;                                                         \
    } while (0)
#define EPV_PROF_ADD(i, v) atomicAdd(&g_solve_prof[i], (unsigned long long)(v))
#else
#define EPV_PROF_BEGIN
#define EPV_PROF(i)
#define EPV_PROF_ADD(i, v)
#endif

__device__ __forceinline__ void null_space_5x9(const double (&x1)[5][2], const double (&x2)[5][2],
                                               double (&e)[4][9]) {
    // A = Q' (9 x 5), column c = epipolar row of correspondence c
    double A[9][5];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        const double a1 = x1[c][0], b1 = x1[c][1], a2 = x2[c][0], b2 = x2[c][1];
        A[0][c] = a2 * a1; A[1][c] = a2 * b1; A[2][c] = a2;
        A[3][c] = b2 * a1; A[4][c] = b2 * b1; A[5][c] = b2;
        A[6][c] = a1;      A[7][c] = b1;      A[8][c] = 1.0;
    }
    double beta[5];
    // Householder QR: reflector k zeroes A[k+1.., k]; v_k is stored in A[k.., k] (v_k[k] = 1 implied)
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        double s = 0.0;
#pragma unroll
        for (int r = k; r < 9; ++r) s += A[r][k] * A[r][k];
        const double nrm = sqrt(s);
        const double alpha = A[k][k] > 0 ? -nrm : nrm;
        const double v0 = A[k][k] - alpha;
        // beta = 2 / (v'v), with v = (v0, A[k+1..][k]);  v'v = s - A[k][k]^2 + v0^2
        const double vtv = s - A[k][k] * A[k][k] + v0 * v0;
        beta[k] = vtv > 0 ? 2.0 / vtv : 0.0;
        A[k][k] = v0;
#pragma unroll
        for (int c = k + 1; c < 5; ++c) {
            double d = 0.0;
#pragma unroll
            for (int r = k; r < 9; ++r) d += A[r][k] * A[r][c];
            d *= beta[k];
#pragma unroll
            for (int r = k; r < 9; ++r) A[r][c] -= d * A[r][k];
        }
    }
    // null space = last four columns of Q_full = H_0 ... H_4 applied to e_5 .. e_8
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        double v[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) v[r] = (r == 5 + b) ? 1.0 : 0.0;
#pragma unroll
        for (int k = 4; k >= 0; --k) {
            double d = 0.0;
#pragma unroll
            for (int r = k; r < 9; ++r) d += A[r][k] * v[r];
            d *= beta[k];
#pragma unroll
            for (int r = k; r < 9; ++r) v[r] -= d * A[r][k];
        }
#pragma unroll
        for (int r = 0; r < 9; ++r) e[b][r] = v[r];
    }
}

// Gauss-Jordan with partial pivoting: left 10x10 -> I, right 10x10 -> inv(left) * right.
__device__ __forceinline__ bool gauss_jordan_10x20(double (&A)[10][20]) {
    for (int k = 0; k < 10; ++k) {
        int p = k;
        double best = fabs(A[k][k]);
        for (int r = k + 1; r < 10; ++r) {
            const double v = fabs(A[r][k]);
            if (v > best) { best = v; p = r; }
        }
        if (!(best > 1e-300)) return false;
        if (p != k) {
            for (int c = k; c < 20; ++c) { const double t = A[k][c]; A[k][c] = A[p][c]; A[p][c] = t; }
        }
        const double inv = 1.0 / A[k][k];
        for (int c = k; c < 20; ++c) A[k][c] *= inv;
        for (int r = 0; r < 10; ++r) {
            if (r == k) continue;
            const double f = A[r][k];
            if (f == 0.0) continue;
            for (int c = k; c < 20; ++c) A[r][c] -= f * A[k][c];
        }
    }
    return true;
}

// ascending-power polynomial product: out[0..na+nb] = a[0..na] * b[0..nb]
template <int NA, int NB>
__device__ __forceinline__ void pmul(const double (&a)[NA + 1], const double (&b)[NB + 1], double (&out)[NA + NB + 1]) {
#pragma unroll
    for (int i = 0; i <= NA + NB; ++i) out[i] = 0.0;
#pragma unroll
    for (int i = 0; i <= NA; ++i)
#pragma unroll
        for (int j = 0; j <= NB; ++j) out[i + j] += a[i] * b[j];
}

// cv::solvePoly's Durand-Kerner on real coefficients c[0..n] (ascending), n <= 10.
// dk_sweeps<N>: degree known at compile time, so the roots live in registers and both loops are
// fully unrolled (the generic version indexes re[]/im[] dynamically, i.e. through local memory).
// Same start values, same Gauss-Seidel sweep order and same stopping rule as the generic version.
template <int N>
__device__ __forceinline__ void dk_sweeps(const double (&c)[11], double (&re)[10], double (&im)[10]) {
    {
        double pr = 1.0, pi = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {                        // roots[i] = (1 + 1i)^i
            re[i] = pr; im[i] = pi;
            const double t = pr - pi;
            pi = pr + pi; pr = t;
        }
    }
    double prev2 = 1e300;
#pragma unroll 1
    for (int iter = 0; iter < 300; ++iter) {
        double maxrel2 = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double xr = re[i], xi = im[i];
            double nr = c[N], ni = 0.0, dr = c[N], di = 0.0;
#pragma unroll
            for (int j = 0; j < N; ++j) {
                const double t = nr * xr - ni * xi + c[N - j - 1];
                ni = nr * xi + ni * xr;
                nr = t;
                if (j != i) {
                    const double er = xr - re[j], ei = xi - im[j];
                    const bool nz = (er != 0.0 || ei != 0.0);
                    const double u = dr * er - di * ei;
                    const double v = dr * ei + di * er;
                    di = nz ? v : di;
                    dr = nz ? u : dr;
                }
            }
            const double s = 1.0 / (dr * dr + di * di);
            const double qr = (nr * dr + ni * di) * s;
            const double qi = (ni * dr - nr * di) * s;
            re[i] = xr - qr; im[i] = xi - qi;
            maxrel2 = fmax(maxrel2, (qr * qr + qi * qi) / fmax(1.0, xr * xr + xi * xi));
        }
        EPV_PROF_ADD(5, 1);
        if (!(maxrel2 > 1e-26)) break;
        if (maxrel2 < 1e-12 && maxrel2 > 1e-2 * prev2) break;
        prev2 = maxrel2;
    }
}

// Generic degree (the leading coefficients vanished): rare, kept out of line.
__device__ __noinline__ void dk_sweeps_generic(const double (&c)[11], int n, double (&re)[10], double (&im)[10]) {
    double pr = 1.0, pi = 0.0;
    for (int i = 0; i < n; ++i) {                            // roots[i] = (1 + 1i)^i
        re[i] = pr; im[i] = pi;
        const double t = pr - pi;
        pi = pr + pi; pr = t;
    }
    double prev2 = 1e300;
    for (int iter = 0; iter < 300; ++iter) {
        double maxrel2 = 0.0;                                // max |update|^2 / max(1, |root|^2) of the sweep
        for (int i = 0; i < n; ++i) {
            const double xr = re[i], xi = im[i];
            double nr = c[n], ni = 0.0, dr = c[n], di = 0.0;
            for (int j = 0; j < n; ++j) {
                // num = num * p + c[n-j-1]
                const double t = nr * xr - ni * xi + c[n - j - 1];
                ni = nr * xi + ni * xr;
                nr = t;
                if (j != i) {
                    const double er = xr - re[j], ei = xi - im[j];
                    if (er != 0.0 || ei != 0.0) {
                        const double u = dr * er - di * ei;
                        di = dr * ei + di * er;
                        dr = u;
                    }
                }
            }
            const double s = 1.0 / (dr * dr + di * di);
            const double qr = (nr * dr + ni * di) * s;
            const double qi = (ni * dr - nr * di) * s;
            re[i] = xr - qr; im[i] = xi - qi;
            maxrel2 = fmax(maxrel2, (qr * qr + qi * qi) / fmax(1.0, xr * xr + xi * xi));
        }
        // cv::solvePoly only stops when a sweep's update is exactly zero, which practically never
        // happens: it runs all 300 sweeps while the roots jitter at round-off level.  Stopping
        // once every update is below 1e-13 of the root leaves the roots equal to OpenCV's up to
        // that jitter (the solutions are refined on the constraints afterwards anyway).
        // Clustered roots never get below their own noise floor (1e-12 .. 1e-9): once the
        // update is small and no longer shrinking by 10x per sweep (the quadratic phase is
        // over), more sweeps only re-draw the noise -- and would stall the whole warp.
        if (!(maxrel2 > 1e-26)) break;
        if (maxrel2 < 1e-12 && maxrel2 > 1e-2 * prev2) break;
        prev2 = maxrel2;
    }
}

// Returns the degree actually solved; roots in (re, im).
__device__ __forceinline__ int durand_kerner(const double (&c)[11], double (&re)[10], double (&im)[10]) {
    int n = 10;
    for (; n > 1; --n)
        if (fabs(c[n]) > 2.220446049250313e-16) break;      // DBL_EPSILON, as cv::solvePoly
    if (n == 10) dk_sweeps<10>(c, re, im);
    else dk_sweeps_generic(c, n, re, im);
    return n;
}

__device__ __forceinline__ double det3(const double (&M)[3][3]) {
    return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
           M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
}

// Unit null vector of a (near) rank-2 3x3 matrix: the largest of the row cross products.
__device__ __forceinline__ void null_vec3(const double (&M)[3][3], double (&v)[3]) {
    double best = -1.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const int r0 = a, r1 = (a + 1) % 3;
        const double cx = M[r0][1] * M[r1][2] - M[r0][2] * M[r1][1];
        const double cy = M[r0][2] * M[r1][0] - M[r0][0] * M[r1][2];
        const double cz = M[r0][0] * M[r1][1] - M[r0][1] * M[r1][0];
        const double n2 = cx * cx + cy * cy + cz * cz;
        if (n2 > best) { best = n2; v[0] = cx; v[1] = cy; v[2] = cz; }
    }
    const double inv = best > 0 ? 1.0 / sqrt(best) : 0.0;
    v[0] *= inv; v[1] *= inv; v[2] *= inv;
}

// ---- refinement (not in OpenCV) -----------------------------------------------------------
// Nister's elimination works in the chart "coefficient of e[3] = 1" of an ARBITRARY null-space
// basis (OpenCV: LAPACK SVD; here: Householder QR).  When a solution has a small e[3]
// component, or the sample is near-degenerate (low parallax), the expanded degree-10
// polynomial loses digits by an amount that differs from basis to basis.  A few Gauss-Newton
// steps on the ten cubic constraints themselves, on the unit sphere of the four basis
// coefficients, make every implementation converge (quadratically) to the same exact
// solutions.  A step is kept only if it lowers the constraint residual.
__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C) {          // C = A B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}
__device__ __forceinline__ void mat3_mul_nt(const double* A, const double* B, double* C) {       // C = A B'
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j * 3] + A[i * 3 + 1] * B[j * 3 + 1] + A[i * 3 + 2] * B[j * 3 + 2];
}
__device__ __forceinline__ void mat3_mul_tn(const double* A, const double* B, double* C) {       // C = A' B
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i] * B[j] + A[3 + i] * B[3 + j] + A[6 + i] * B[6 + j];
}

__device__ __forceinline__ double constraints10(const double* E, double* F) {
    double G[9], T[9];
    mat3_mul_nt(E, E, G);                      // E E'
    mat3_mul(G, E, T);
    const double tr = G[0] + G[4] + G[8];
    double m = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) { F[i] = 2.0 * T[i] - tr * E[i]; m = fmax(m, fabs(F[i])); }
    F[9] = E[0] * (E[4] * E[8] - E[5] * E[7]) - E[1] * (E[3] * E[8] - E[5] * E[6]) + E[2] * (E[3] * E[7] - E[4] * E[6]);
    return fmax(m, fabs(F[9]));
}

__device__ __noinline__ void refine_essential(const double (&e)[4][9], double (&E)[9]) {
    double c[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 9; ++i) s += E[i] * e[k][i];
        c[k] = s;
    }
    {
        const double n = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]);
        if (!(n > 0)) return;
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] /= n;
    }
    double Ec[9], F[10];
#pragma unroll
    for (int i = 0; i < 9; ++i) Ec[i] = c[0] * e[0][i] + c[1] * e[1][i] + c[2] * e[2][i] + c[3] * e[3][i];
    double fmaxv = constraints10(Ec, F);
    for (int it = 0; it < 6; ++it) {
        // Jacobian columns: dF along each basis matrix
        double J[10][4];
        double G[9], EtE[9], cof[9];
        mat3_mul_nt(Ec, Ec, G);                 // E E'
        mat3_mul_tn(Ec, Ec, EtE);               // E'E
        const double tr = G[0] + G[4] + G[8];
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                cof[i * 3 + j] = Ec[i1 * 3 + j1] * Ec[i2 * 3 + j2] - Ec[i1 * 3 + j2] * Ec[i2 * 3 + j1];
            }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double* D = e[k];
            double A1[9], A2[9], A3[9], T[9];
            mat3_mul(D, EtE, A1);               // D E'E
            mat3_mul_nt(Ec, D, T);              // E D'
            mat3_mul(T, Ec, A2);                // E D' E
            mat3_mul(G, D, A3);                 // E E' D
            const double trED = T[0] + T[4] + T[8];
            double dd = 0.0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                J[i][k] = 2.0 * (A1[i] + A2[i] + A3[i]) - 2.0 * trED * Ec[i] - tr * D[i];
                dd += cof[i] * D[i];
            }
            J[9][k] = dd;
        }
        double N[4][5];
        double trN = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < 10; ++r) s += J[r][a] * J[r][b];
                N[a][b] = s;
            }
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < 10; ++r) s += J[r][a] * F[r];
            N[a][4] = -s;
            trN += N[a][a];
        }
        const double mu = 1e3 * trN;             // pins the radial (scale) direction
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) N[a][b] += mu * c[a] * c[b];
        // 4x4 Gaussian elimination with partial pivoting
        bool ok = true;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int p = k;
            double best = fabs(N[k][k]);
#pragma unroll
            for (int r = k + 1; r < 4; ++r)
                if (fabs(N[r][k]) > best) { best = fabs(N[r][k]); p = r; }
            if (!(best > 0)) { ok = false; break; }
#pragma unroll
            for (int r = k + 1; r < 4; ++r)
                if (r == p) {
#pragma unroll
                    for (int q = 0; q < 5; ++q) { const double t = N[k][q]; N[k][q] = N[r][q]; N[r][q] = t; }
                }
#pragma unroll
            for (int r = k + 1; r < 4; ++r) {
                const double f = N[r][k] / N[k][k];
#pragma unroll
                for (int q = k; q < 5; ++q) N[r][q] -= f * N[k][q];
            }
        }
        if (!ok) break;
        double d[4];
#pragma unroll
        for (int r = 3; r >= 0; --r) {
            double s = N[r][4];
#pragma unroll
            for (int q = r + 1; q < 4; ++q) s -= N[r][q] * d[q];
            d[r] = s / N[r][r];
        }
        double c2[4], n2 = 0.0, dn = 0.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { c2[k] = c[k] + d[k]; n2 += c2[k] * c2[k]; dn += d[k] * d[k]; }
        n2 = sqrt(n2);
        double E2[9], F2[10];
#pragma unroll
        for (int k = 0; k < 4; ++k) c2[k] /= n2;
#pragma unroll
        for (int i = 0; i < 9; ++i) E2[i] = c2[0] * e[0][i] + c2[1] * e[1][i] + c2[2] * e[2][i] + c2[3] * e[3][i];
        const double f2 = constraints10(E2, F2);
        if (!(f2 < fmaxv)) break;                // NaN or no improvement: keep the current iterate
        fmaxv = f2;
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = c2[k];
#pragma unroll
        for (int i = 0; i < 9; ++i) Ec[i] = E2[i];
#pragma unroll
        for (int i = 0; i < 10; ++i) F[i] = F2[i];
        if (sqrt(dn) < 1e-14) break;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) E[i] = Ec[i];
}

// Solve one sample.  Eout[k] = k-th essential matrix (row-major, unit Frobenius norm).
__device__ __noinline__ int solve(const double (&x1)[5][2], const double (&x2)[5][2], double (*Eout)[9]) {
    EPV_PROF_BEGIN;
    double e[4][9];
    null_space_5x9(x1, x2, e);
    double A[10][20];
    fivept_constraints(e, A);
    EPV_PROF(0);
    if (!gauss_jordan_10x20(A)) return 0;
    EPV_PROF(1);
    // B(z): entries (j,0),(j,1) cubic, (j,2) quartic; ascending powers.  Row j comes from
    // reduced rows 4+2j ("e - z f"): coefficient layout of the right block per row is
    // [xz^2 xz x | yz^2 yz y | z^3 z^2 z 1] (descending in z inside each group).
    double B[3][3][5];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double* r1 = &A[4 + 2 * j][10];
        const double* r2 = &A[5 + 2 * j][10];
        // group g (x: 0..2, y: 3..5) -> cubic: r1 contributes z^2..z^0, -z*r2 contributes z^3..z^1
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const double* a = r1 + 3 * g;
            const double* b = r2 + 3 * g;
            B[j][g][0] = a[2];
            B[j][g][1] = a[1] - b[2];
            B[j][g][2] = a[0] - b[1];
            B[j][g][3] = -b[0];
            B[j][g][4] = 0.0;
        }
        const double* a = r1 + 6;
        const double* b = r2 + 6;
        B[j][2][0] = a[3];
        B[j][2][1] = a[2] - b[3];
        B[j][2][2] = a[1] - b[2];
        B[j][2][3] = a[0] - b[1];
        B[j][2][4] = -b[0];
    }
    double c[11];
    {
        double c3a[4], c3b[4], q4[5], m7a[8], m7b[8], m6a[7], m6b[7], t10[11];
#pragma unroll
        for (int i = 0; i < 11; ++i) c[i] = 0.0;
        // + B00 * (B11*B22 - B12*B21)
#pragma unroll
        for (int i = 0; i < 4; ++i) { c3a[i] = B[1][1][i]; c3b[i] = B[2][1][i]; }
#pragma unroll
        for (int i = 0; i < 5; ++i) q4[i] = B[2][2][i];
        pmul<3, 4>(c3a, q4, m7a);
#pragma unroll
        for (int i = 0; i < 5; ++i) q4[i] = B[1][2][i];
        pmul<3, 4>(c3b, q4, m7b);
#pragma unroll
        for (int i = 0; i < 8; ++i) m7a[i] -= m7b[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) c3a[i] = B[0][0][i];
        pmul<3, 7>(c3a, m7a, t10);
#pragma unroll
        for (int i = 0; i < 11; ++i) c[i] += t10[i];
        // - B01 * (B10*B22 - B12*B20)
#pragma unroll
        for (int i = 0; i < 4; ++i) { c3a[i] = B[1][0][i]; c3b[i] = B[2][0][i]; }
#pragma unroll
        for (int i = 0; i < 5; ++i) q4[i] = B[2][2][i];
        pmul<3, 4>(c3a, q4, m7a);
#pragma unroll
        for (int i = 0; i < 5; ++i) q4[i] = B[1][2][i];
        pmul<3, 4>(c3b, q4, m7b);
#pragma unroll
        for (int i = 0; i < 8; ++i) m7a[i] -= m7b[i];
#pragma unroll
        for (int i = 0; i < 4; ++i) c3a[i] = B[0][1][i];
        pmul<3, 7>(c3a, m7a, t10);
#pragma unroll
        for (int i = 0; i < 11; ++i) c[i] -= t10[i];
        // + B02 * (B10*B21 - B11*B20)
#pragma unroll
        for (int i = 0; i < 4; ++i) { c3a[i] = B[1][0][i]; c3b[i] = B[2][1][i]; }
        pmul<3, 3>(c3a, c3b, m6a);
#pragma unroll
        for (int i = 0; i < 4; ++i) { c3a[i] = B[1][1][i]; c3b[i] = B[2][0][i]; }
        pmul<3, 3>(c3a, c3b, m6b);
#pragma unroll
        for (int i = 0; i < 7; ++i) m6a[i] -= m6b[i];
#pragma unroll
        for (int i = 0; i < 5; ++i) q4[i] = B[0][2][i];
        pmul<4, 6>(q4, m6a, t10);
#pragma unroll
        for (int i = 0; i < 11; ++i) c[i] += t10[i];
    }
    EPV_PROF(2);
    double re[10], im[10];
    const int n = durand_kerner(c, re, im);
    EPV_PROF(3);
    int count = 0;
    for (int i = 0; i < n; ++i) {
        if (fabs(im[i]) > 1e-10) continue;
        const double z = re[i];
        double Bz[3][3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int g = 0; g < 3; ++g) {
                double acc = B[j][g][4];
#pragma unroll
                for (int k = 3; k >= 0; --k) acc = acc * z + B[j][g][k];
                Bz[j][g] = acc;
            }
        }
        double v[3];
        null_vec3(Bz, v);
        if (!(fabs(v[2]) >= 1e-10)) continue;
        const double x = v[0] / v[2], y = v[1] / v[2];
        double s = 0.0;
        double E[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            E[k] = e[0][k] * x + e[1][k] * y + e[2][k] * z + e[3][k];
            s += E[k] * E[k];
        }
        const double inv = 1.0 / sqrt(s);
#pragma unroll
        for (int k = 0; k < 9; ++k) E[k] *= inv;
        refine_essential(e, E);
#pragma unroll
        for (int k = 0; k < 9; ++k) Eout[count][k] = E[k];
        ++count;
    }
    EPV_PROF(4);
    EPV_PROF_ADD(6, 1);
    EPV_PROF_ADD(7, count);
    return count;
}

}  // namespace fivept
